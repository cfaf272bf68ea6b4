#!/bin/bash
# forward warp kernel: L2 prefetch distance in CTAs (0 = off)
for ah in 0 600 1200 2400 4800; do for w in cfg2 cfg1; do
  MT_WARP_AHEAD=$ah timeout 120 python bench.py --workload $w --steps 300 --warmup 5 --no-cpu-baseline --e2e-steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ahead=$ah $w step=%.1f us  '%(d['ms_per_step']*1e3) + ' '.join('%s=%.1f(%.2f)'%(k['call'][3:],k['avg_us'],k['frac_hbm']) for k in d['kernels']))"
done; done
