#!/bin/bash
# staged (smem + cp.async.bulk) vs direct-gather forward warp kernel, CPN (cfg2) and DFPN (cfg1) paths
for st in 0 1; do for w in cfg2 cfg1 cfg4; do
  MT_WARP_STAGED=$st timeout 120 python bench.py --workload $w --steps 300 --warmup 5 --no-cpu-baseline --e2e-steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('staged=$st $w  step=%.1f us  '%(d['ms_per_step']*1e3) + '  '.join('%s=%.1f(%.2f)'%(k['call'][3:],k['avg_us'],k['frac_hbm']) for k in d['kernels']))"
done; done
