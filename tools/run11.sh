timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
run() { env "$@" timeout 180 python bench.py --workload ${WL:-cfg2} --steps 300 --warmup 5 --no-cpu-baseline --e2e-steps 2 2>gpurun_out/err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('${WL:-cfg2} $*  step=%.1f us  '%(d['ms_per_step']*1e3) + '  '.join('%s=%.1f'%(k['call'][3:],k['avg_us']) for k in d['kernels']))" || tail -5 gpurun_out/err.log; }
run MT_X=0
run MT_PDL=0
run MT_CM_SIM_CH=4
run MT_CM_COPY_CH=2
run MT_CM_SIM_CH=4 MT_CM_COPY_CH=2
B=32; run2() { env "$@" timeout 180 python bench.py --workload cfg2 --batch $B --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 1 2>gpurun_out/err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('B=$B $*  step=%.1f us  '%(d['ms_per_step']*1e3) + '  '.join('%s=%.1f'%(k['call'][3:],k['avg_us']) for k in d['kernels'][:2]))" || tail -5 gpurun_out/err.log; }
run2 MT_X=0
