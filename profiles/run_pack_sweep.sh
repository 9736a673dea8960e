FLAGS="--e2e-steps 0 --no-cpu-baseline --no-backbone-view --rle-steps 0 --steps 50 --warmup 5"
pick='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_pass"], d["ms_per_pass_serial"], {k:(v["ms"],v["ms_alone"]) for k,v in d["kernels"].items()})'
for v in 1 2 3 4; do echo "== pack ctas/sm $v"; HGL_LIB=$PWD/hybridgl_b200/libhgl_tuning.so HGL_PACK_CTAS_PER_SM=$v timeout 300 python bench.py $FLAGS 2>/dev/null | python -c "$pick"; done
for v in 2 6; do echo "== rows ctas/sm $v"; HGL_LIB=$PWD/hybridgl_b200/libhgl_tuning.so HGL_ROWS_CTAS_PER_SM=$v timeout 300 python bench.py $FLAGS 2>/dev/null | python -c "$pick"; done
