#!/bin/bash
# device-resident bench line (no e2e / cpu / backbone / rle legs) for the product library and every variant library
FLAGS="--e2e-steps 0 --no-cpu-baseline --no-backbone-view --rle-steps 0 --steps 50 --warmup 5"
pick='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_pass"], d["ms_per_pass_serial"], {k:(v["ms"],v["ms_alone"]) for k,v in d["kernels"].items() if k in ("prep","grid_heat_pool","pack","blur")})'
echo "== product"; timeout 300 python bench.py $FLAGS "$@" 2>/dev/null | python -c "$pick"
for lib in profiles/_variants/*.so; do
  echo "== $lib"; HGL_LIB=$PWD/$lib timeout 300 python bench.py $FLAGS "$@" 2>/dev/null | python -c "$pick"
done
