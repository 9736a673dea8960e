#!/bin/bash
# The other BASELINE.json configs through bench.py (not bench lines: parity cases; this records their step times).
out=gpurun_out; mkdir -p $out
for w in 1 3 4 5; do
  python bench.py --workload $w --images 16 --steps 100 --warmup 5 --e2e-steps 3 --rle-steps 20 --serial-steps 20 --no-cpu-baseline 2>$out/cfg$w.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels']; c=d['config']
print('| configs[%d] | %dx%d, %d masks, %d expr, S=%d g=%d De=%d | %.3f | %.0f | %.0f | %.3f | %.0f | %s |' % ($w-1, c['h'],c['w'],c['n_masks'],c['n_expr'],c['S'],c['g'],c['De'], d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac_alone'], d['rle_input']['e2e']['value'], ' '.join('%s %.0f'%(n,1e3*k[n].get('ms_alone',0)) for n in k)))"
  tail -n 2 $out/cfg$w.err
done
