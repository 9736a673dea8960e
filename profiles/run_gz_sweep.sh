#!/bin/bash
# prep main: z-slices per band tile (masks per CTA = masks per image / gz), profiling build
FLAGS="--e2e-steps 0 --no-cpu-baseline --no-backbone-view --rle-steps 0 --steps 50 --warmup 5"
pick='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_pass"], d["ms_per_pass_serial"], {k:(v["ms"],v["ms_alone"]) for k,v in d["kernels"].items() if k in ("prep","grid_heat_pool")})'
for v in ${@:-1 2 3 4 5 6 7 9 12}; do echo "== HGL_PREP_GZ=$v"; HGL_LIB=$PWD/hybridgl_b200/libhgl_tuning.so HGL_PREP_GZ=$v timeout 300 python bench.py $FLAGS 2>/dev/null | python -c "$pick"; done
