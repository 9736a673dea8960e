FLAGS="--e2e-steps 0 --no-cpu-baseline --no-backbone-view --rle-steps 0 --steps 50 --warmup 5"
pick='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_pass"], d["ms_per_pass_serial"], d["roofline"]["frac"], {k:(v["ms"],v["ms_alone"]) for k,v in d["kernels"].items()}); print(d.get("timeline_ms",{}).get("overlapped"))'
echo "== default"; timeout 300 python bench.py $FLAGS 2>/dev/null | python -c "$pick"
echo "== prefetch"; timeout 300 python bench.py $FLAGS --prefetch 2>/dev/null | python -c "$pick"
echo "== rows-first"; timeout 300 python bench.py $FLAGS --rows-first 1 2>/dev/null | python -c "$pick"
echo "== prefetch + rows-first"; timeout 300 python bench.py $FLAGS --prefetch --rows-first 1 2>/dev/null | python -c "$pick"
