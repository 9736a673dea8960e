#!/bin/bash
# One gpurun --gpus N call: the torchrun launches the driver uses (bench, reference arm).
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash profiles/run_multi.sh r2 2'
tag=${1:-multi}; n=${2:-2}
out=gpurun_out; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29512 bench.py --gpus $n --steps 20 --warmup 5 > $out/${tag}_bench${n}.json 2> $out/${tag}_bench${n}.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('$out/${tag}_bench${n}.json')); print('N', d['n_gpus'], 'value', round(d['value']), 'ms/pass', round(d['ms_per_pass'],4), 'coll ms', d['collective_ms'], 'e2e', round(d['e2e']['value']), 'rle e2e', round(d['rle_input']['e2e']['value']))"
timeout 400 $TR --master-port 29513 bench.py --impl reference --gpus $n --steps 20 --warmup 5 > $out/${tag}_ref${n}.json 2> $out/${tag}_ref${n}.err; echo "ref rc=$?"
cut -c1-700 $out/${tag}_ref${n}.json
tail -n 2 $out/${tag}_bench${n}.err $out/${tag}_ref${n}.err
