set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
run() { env "$@" timeout 180 python bench.py --workload ${WL:-cfg2} --steps 300 --warmup 5 --no-cpu-baseline --e2e-steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('${WL:-cfg2} $*  step=%.1f us  '%(d['ms_per_step']*1e3) + '  '.join('%s=%.1f'%(k['call'][3:],k['avg_us']) for k in d['kernels']))"; }
run MT_B200_LIB=$PWD/tools/libmt_map32.so MT_CM_FUSED=0
run MT_CM_FUSED=0
run MT_CM_FUSED=1
run MT_CM_FUSED=1 MT_CM_FUSED_MINB=2
run MT_CM_FUSED=1 MT_CM_FUSED_CH=4
run MT_CM_FUSED=1 MT_CM_LAG=2
run MT_CM_FUSED=1 MT_CM_LAG=2 MT_CM_FUSED_MINB=2
WL=cfg5 run MT_CM_FUSED=0
WL=cfg5 run MT_CM_FUSED=1
