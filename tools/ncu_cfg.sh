#!/bin/bash
# ncu --set full of one launch of every distinct kernel of a workload: tools/ncu_cfg.sh cfg3
WL=$1
CMD="python bench.py --workload $WL --steps 12 --warmup 3 --no-cpu-baseline --e2e-steps 1"
$CMD > gpurun_out/plain_$WL.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"corr|warp_l1|chn_|masked|hole|warp_fwd" -s 40 -c ${2:-8} -f -o gpurun_out/prof_$WL $CMD > gpurun_out/ncu_$WL.log 2>&1
tail -2 gpurun_out/ncu_$WL.log
