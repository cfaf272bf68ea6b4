#!/bin/bash
# forward warp kernel: rows per iteration x row groups per thread
for rows in 2 4; do for it in 1 2 4 8; do
  MT_WARP_ROWS=$rows MT_WARP_ITERS=$it timeout 120 python bench.py --workload cfg2 --steps 300 --warmup 5 --no-cpu-baseline --e2e-steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('rows=$rows iters=$it  step=%.1f us  '%(d['ms_per_step']*1e3) + '  '.join('%s=%.1f'%(k['call'][3:],k['avg_us']) for k in d['kernels']))"
done; done
