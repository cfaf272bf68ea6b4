timeout 900 python -m pytest tests -m gpu -x -q -k "cm or full_size" 2>&1 | tail -3
run() { env "$@" timeout 180 python bench.py --workload cfg2 ${B:+--batch $B} --steps 300 --warmup 5 --no-cpu-baseline --e2e-steps 2 2>gpurun_out/err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('B=${B:-8} $*  step=%.1f us  '%(d['ms_per_step']*1e3) + '  '.join('%s=%.1f'%(k['call'][3:],k['avg_us']) for k in d['kernels']))" || tail -5 gpurun_out/err.log; }
run MT_CM_TABLE=0
run MT_CM_TABLE=1
run MT_CM_TABLE=1 MT_CM_COPY_CH=2
run MT_CM_TABLE=1 MT_CM_SIM_CH=2
run MT_CM_TABLE=1 MT_CM_COPY_REVERSE=0
B=32
run MT_CM_TABLE=0
run MT_CM_TABLE=1
run MT_CM_TABLE=1 MT_CM_COPY_CH=2
