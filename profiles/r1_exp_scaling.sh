run() { env "$@" timeout 180 python bench.py --workload cfg2 --batch $B --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('B=$B $*  step=%.1f us  '%(d['ms_per_step']*1e3) + '  '.join('%s=%.1f'%(k['call'][3:],k['avg_us']) for k in d['kernels'][:2]))"; }
for B in 8 16 32; do
run MT_WARP_STAGED=0
run MT_WARP_STAGED=1
run MT_WARP_STAGED=1 MT_BENCH_THETA_SIGMA=0.0
run MT_WARP_STAGED=0 MT_BENCH_THETA_SIGMA=0.0
done
