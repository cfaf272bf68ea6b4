#!/bin/bash
# CM chunk sweep (samples per sim->weights->copy group); 8 = whole batch at once
for ch in 1 2 4 8; do
  MT_CM_CHUNK=$ch timeout 120 python bench.py --workload cfg2 --steps 300 --warmup 5 --no-cpu-baseline --e2e-steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('cm_chunk=$ch  step=%.1f us  '%(d['ms_per_step']*1e3) + '  '.join('%s=%.1f'%(k['call'][3:],k['avg_us']) for k in d['kernels']))"
done
