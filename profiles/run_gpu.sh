#!/bin/bash
# One gpurun call = parity tests + bench + ncu launch list + ncu --set full of our kernels (1 GPU).
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash profiles/run_gpu.sh r1a'
# Outputs land in gpurun_out/<tag>_*; summaries are copied into profiles/ by profiles/summarize.py (run here, no GPU).
tag=${1:-run}
out=gpurun_out
KRE='prep_main|prep_setup|prep_crop|pack_masks|blur15|mask_grid|mask_area|mask_rows|mask_geometry|heat_prefix|heat_consts|heat_resize|score_select|pool_score|iou_kernel|iou_zero|token_mask|attn_|rle_|gem_|cls_'
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,memory.total --format=csv > $out/${tag}_gpu.csv 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> $out/${tag}_smoke.log
tail -2 $out/${tag}_smoke.log
timeout 400 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
cat $out/${tag}_bench.json
# launch list (cold-cache, serialised: shares only)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:$KRE" -c 60 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-graph --inner 1 --e2e-steps 0 --no-cpu-baseline --no-backbone-view --serial-steps 0 --rle-steps 0 > $out/${tag}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
# full sections for one step of our kernels (second pass over the path: skip the first step's launches)
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$KRE" -s ${NCU_SKIP:-10} -c ${NCU_COUNT:-10} -o $out/${tag}_prof -f \
    python bench.py --steps 1 --warmup 1 --no-graph --inner 1 --e2e-steps 0 --no-cpu-baseline --no-backbone-view --serial-steps 0 --rle-steps 0 > $out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $out | tail -12
