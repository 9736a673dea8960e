#!/bin/bash
# bench.py under torchrun at N GPUs, short form (scaling check):  gpurun --gpus 8 -- 'bash profiles/run_scale.sh r1y 8'
tag=${1:-scale}; n=${2:-8}
out=gpurun_out; mkdir -p $out
nproc; free -g | head -2
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $n --steps 200 --warmup 10 --serial-steps 0 --rle-steps 20 --e2e-steps 5 > $out/${tag}_bench${n}.json 2> $out/${tag}_bench${n}.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('$out/${tag}_bench${n}.json')); print('N', d['n_gpus'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'rle', round(d['rle_input']['value']), 'rle e2e', round(d['rle_input']['e2e']['value']), d['clocks'])"
tail -n 3 $out/${tag}_bench${n}.err
