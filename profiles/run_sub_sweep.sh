#!/bin/bash
# prep main: masks per ring stage (HGL_PREP_SUB) x masks between two outline flushes (HGL_PREP_FLUSH), profiling build
#   bash profiles/run_sub_sweep.sh "4:4 4:16 8:8 8:32"
FLAGS="--e2e-steps 0 --no-cpu-baseline --no-backbone-view --rle-steps 0 --steps 40 --warmup 5"
pick='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_pass"], d["ms_per_pass_serial"], {k:(v["ms"],v["ms_alone"]) for k,v in d["kernels"].items() if k in ("prep",)})'
for pair in ${1:-2:1 2:2 4:1 4:2 4:4 8:1 8:2 8:4 8:8}; do sub=${pair%%:*}; fl=${pair##*:}; echo "== SUB=$sub FLUSH=$fl"; HGL_LIB=$PWD/hybridgl_b200/libhgl_tuning.so HGL_PREP_SUB=$sub HGL_PREP_FLUSH=$fl timeout 300 python bench.py $FLAGS 2>/dev/null | python -c "$pick"; done
