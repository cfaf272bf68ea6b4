#!/bin/bash
# warp kernels: __launch_bounds__(128, MINB) builds (python -m master_thesis_b200.build -DMT_WARP_MINB=n
# -DMT_WARPB_MINB=n --out=master_thesis_b200/sweep_mb<n>.so), selected with MT_B200_LIB
for wl in cfg2 cfg1 cfg3 cfg4; do for mb in 1 4 6 8 12; do
  lib=/root/repo/master_thesis_b200/sweep_mb$mb.so; [ $mb = 6 ] && lib=/root/repo/master_thesis_b200/libmt_b200.so
  [ -f $lib ] || continue
  MT_B200_LIB=$lib timeout 180 python bench.py --workload $wl --steps 300 --warmup 5 --no-cpu-baseline --e2e-steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$wl minb=$mb  step=%.1f us  '%(d['ms_per_step']*1e3) + '  '.join('%s=%.1f'%(k['call'][3:],k['avg_us']) for k in d['kernels']))"
done; done
