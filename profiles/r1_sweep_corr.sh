#!/bin/bash
# correlation: columns per CTA (256 = 1 CTA per frame, 128 = 2, 64 = 4)
for tn in 256 128 64; do for w in cfg1 cfg3; do
  MT_CORR_TN=$tn timeout 120 python bench.py --workload $w --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('tn=$tn $w step=%.1f us  '%(d['ms_per_step']*1e3) + ' '.join('%s=%.1f'%(k['call'][3:],k['avg_us']) for k in d['kernels'][:2]))"
done; done
