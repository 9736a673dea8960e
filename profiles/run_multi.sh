#!/bin/bash
# One gpurun --gpus N call: the torchrun launches the driver uses (bench, reference arm) + the NCCL sweep check.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash profiles/run_multi.sh r1x 2'
tag=${1:-multi}; n=${2:-2}
out=gpurun_out; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $out/${tag}_gpus${n}.csv 2>&1
timeout 300 $TR --master-port 29511 profiles/sweep_nccl.py --images 24 --per-step 4 > $out/${tag}_sweep${n}.json 2> $out/${tag}_sweep${n}.err; echo "sweep rc=$?"
cat $out/${tag}_sweep${n}.json
timeout 400 $TR --master-port 29512 bench.py --gpus $n --steps 100 --warmup 5 --serial-steps 0 > $out/${tag}_bench${n}.json 2> $out/${tag}_bench${n}.err; echo "bench rc=$?"
cut -c1-900 $out/${tag}_bench${n}.json
timeout 300 $TR --master-port 29513 bench.py --impl reference --gpus $n --steps 2 --warmup 1 > $out/${tag}_ref${n}.json 2> $out/${tag}_ref${n}.err; echo "ref rc=$?"
cut -c1-600 $out/${tag}_ref${n}.json
tail -n 3 $out/${tag}_sweep${n}.err $out/${tag}_bench${n}.err $out/${tag}_ref${n}.err
