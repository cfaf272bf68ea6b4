run2() { env "$@" timeout 180 python bench.py --workload cfg2 --batch $B --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 1 2>gpurun_out/err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('B=$B $*  step=%.1f us  '%(d['ms_per_step']*1e3) + '  '.join('%s=%.1f'%(k['call'][3:],k['avg_us']) for k in d['kernels'][:2]))" || tail -5 gpurun_out/err.log; }
for B in 16 32 64; do
run2 MT_CM_FUSED=0
run2 MT_CM_FUSED=1
run2 MT_CM_FUSED=1 MT_CM_LAG=4
run2 MT_CM_FUSED=1 MT_CM_LAG=8
run2 MT_CM_FUSED=1 MT_CM_LAG=4 MT_CM_STAGES=3
done
