#!/bin/bash
# round 2, call AB: register budget of the dense-flow loss kernels (MT_WARPB_MINB 6 / 7 / 8)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=$PWD/master_thesis_b200
for v in base wb7 wb8; do
  lib=$L/libmt_$v.so; [ $v = base ] && lib=$L/libmt_b200.so
  MT_B200_LIB=$lib python tools/gpu/probe_cfg3_step.py 2>&1 | grep -E "lib|warp_l1|0\.\.12" | sed "s/^/$v /"
done | tee gpurun_out/ab_probe.txt
MT_B200_LIB=$L/libmt_wb7.so timeout 300 python -m pytest tests -m gpu -q -x -k "loss or dfpn" 2>&1 | tail -2
