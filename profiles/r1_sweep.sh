#!/bin/bash
# tuning sweep: prints per-call microseconds for each knob combination
for rows in 2; do for sc in 2 4; do for cc in 2 4; do
  MT_WARP_ROWS=$rows MT_CM_SIM_CH=$sc MT_CM_COPY_CH=$cc timeout 120 python bench.py --workload cfg2 --steps 300 --warmup 5 --no-cpu-baseline --e2e-steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('rows=$rows sim_ch=$sc copy_ch=$cc  step=%.1f us  '%(d['ms_per_step']*1e3) + '  '.join('%s=%.1f'%(k['call'][3:],k['avg_us']) for k in d['kernels']))"
done; done; done
