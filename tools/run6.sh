set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
MT_CM_FUSED=1 python tools/dbg_cm.py 8 2>&1 | tail -4
run() { env "$@" timeout 180 python bench.py --workload ${WL:-cfg2} --steps 300 --warmup 5 --no-cpu-baseline --e2e-steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('${WL:-cfg2} $*  step=%.1f us  '%(d['ms_per_step']*1e3) + '  '.join('%s=%.1f'%(k['call'][3:],k['avg_us']) for k in d['kernels']))"; }
run MT_CM_FUSED=0 MT_WARP_TILE_W=32
run MT_CM_FUSED=0 MT_WARP_TILE_W=64
run MT_CM_FUSED=1
run MT_CM_FUSED=1 MT_CM_STAGES=4
run MT_CM_FUSED=1 MT_CM_STAGES=3
run MT_CM_FUSED=1 MT_CM_FUSED_CH=1
run MT_CM_FUSED=1 MT_CM_FUSED_CH=4
run MT_CM_FUSED=1 MT_CM_LAG=1
run MT_CM_FUSED=1 MT_CM_LAG=2
run MT_CM_FUSED=1 MT_CM_LAG=4
B=32; run2() { env "$@" timeout 180 python bench.py --workload cfg2 --batch $B --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('B=$B $*  step=%.1f us  '%(d['ms_per_step']*1e3) + '  '.join('%s=%.1f'%(k['call'][3:],k['avg_us']) for k in d['kernels'][:2]))"; }
run2 MT_CM_FUSED=0 MT_WARP_TILE_W=32
run2 MT_CM_FUSED=1 MT_WARP_TILE_W=64
