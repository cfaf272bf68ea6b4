timeout 900 python -m pytest tests -m gpu -x -q -k "loss or full_size or smoke" 2>&1 | tail -3
run() { env "$@" timeout 180 python bench.py --workload ${WL:-cfg2} --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 2 2>gpurun_out/err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('${WL:-cfg2} $*  step=%.1f us  '%(d['ms_per_step']*1e3) + '  '.join('%s=%.1f(%.2f)'%(k['call'][3:],k['avg_us'],k['frac_hbm']) for k in d['kernels']))" || tail -5 gpurun_out/err.log; }
WL=cfg3 run MT_X=1
WL=cfg3 run MT_WARPL1_CTAS_PER_SM=12
WL=cfg3 run MT_WARPL1_CTAS_PER_SM=8
