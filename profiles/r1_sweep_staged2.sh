#!/bin/bash
# persistent TMA-staged CPN warp kernel (warp_tma.cu) vs the direct-gather kernel, cfg2 / cfg5
run() { env "$@" timeout 180 python bench.py --workload $WL --steps 300 --warmup 5 --no-cpu-baseline --e2e-steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$WL $*  step=%.1f us  '%(d['ms_per_step']*1e3) + '  '.join('%s=%.1f'%(k['call'][3:],k['avg_us']) for k in d['kernels']))"; }
for WL in cfg2 cfg5; do
run MT_WARP_STAGED=0
run MT_WARP_STAGED=1 MT_WARP_BOX=48
run MT_WARP_STAGED=1 MT_WARP_BOX=40
done
