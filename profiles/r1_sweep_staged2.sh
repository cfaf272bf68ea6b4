#!/bin/bash
# persistent TMA-staged CPN warp kernel (warp_tma.cu) vs the direct-gather kernel
run() { env "$@" timeout 180 python bench.py --workload $WL --steps 300 --warmup 5 --no-cpu-baseline --e2e-steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$WL $*  step=%.1f us  '%(d['ms_per_step']*1e3) + '  '.join('%s=%.1f'%(k['call'][3:],k['avg_us']) for k in d['kernels'][:2]))"; }
for WL in ${WLS:-cfg2 cfg5}; do
run MT_WARP_STAGED=0
for cps in 1 2; do for box in 48 40; do
run MT_WARP_STAGED=1 MT_WARP_BOX=$box MT_WARP_CTAS_PER_SM=$cps
done; done; done
