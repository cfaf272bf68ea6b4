#!/bin/bash
# CM variants at cfg2 (B=8) and B=32: three launches (default), merged copy+sim groups, the
# experimental persistent pipelined kernel, slab widths, PDL on/off
run() { env "$@" timeout 180 python bench.py --workload cfg2 ${B:+--batch $B} --steps 300 --warmup 5 --no-cpu-baseline --e2e-steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('B=${B:-8} $*  step=%.1f us  '%(d['ms_per_step']*1e3) + '  '.join('%s=%.1f'%(k['call'][3:],k['avg_us']) for k in d['kernels']))"; }
for B in "" 32; do
run MT_X=0
run MT_PDL=0
run MT_CM_SIM_CH=2
run MT_CM_COPY_CH=2
run MT_CM_COPY_REVERSE=0
run MT_CM_CHUNK=4
run MT_CM_CHUNK=2
run MT_CM_FUSED=1
run MT_CM_FUSED=1 MT_CM_LAG=4
run MT_WARP_TILE_W=64
run MT_WARP_STAGED=0
done
