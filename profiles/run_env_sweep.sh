#!/bin/bash
# bench under a sweep of one tuning environment variable:  bash profiles/run_env_sweep.sh VAR v1 v2 ...
var=$1; shift
for v in "$@"; do
  env $var=$v python bench.py --steps 100 --warmup 5 --e2e-steps 0 --no-cpu-baseline --rle-steps 0 --serial-steps 30 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=d['kernels']
print('$var=$v', 'ms/step %.4f'%d['ms_per_step'], 'serial %.4f'%d['ms_per_step_serial'], ' '.join('%s %.0f/%.0f'%(n,1e3*k[n]['ms'],1e3*k[n].get('ms_alone',0)) for n in k))"
done
