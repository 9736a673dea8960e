#!/bin/bash
# bench.py under torchrun at N GPUs, the driver's command line (scaling check):  gpurun --gpus 8 -- 'bash profiles/run_scale.sh r2 8'
tag=${1:-scale}; n=${2:-8}
out=gpurun_out; mkdir -p $out
nproc; free -g | head -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $n --steps 200 --warmup 10 > $out/${tag}_bench${n}.json 2> $out/${tag}_bench${n}.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('$out/${tag}_bench${n}.json')); print('N', d['n_gpus'], 'value', round(d['value']), 'ms/pass', round(d['ms_per_pass'],4), 'e2e', round(d['e2e']['value']), d['e2e'].get('link_gbs_per_gpu'), 'rle', round(d['rle_input']['value']), 'rle e2e', round(d['rle_input']['e2e']['value']), d['rle_input']['e2e'].get('link_gbs_per_gpu'), d['clocks'])"
tail -n 3 $out/${tag}_bench${n}.err
