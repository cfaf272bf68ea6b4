#!/bin/bash
# round 2, call AC: CTA cap of the fused-loss forward (waves of resident CTAs), backward at 64 registers
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in 16 6 12 24 31; do
  MT_WARPL1_CTAS_PER_SM=$v python tools/gpu/probe_cfg3_step.py 2>&1 | grep -E "warp_l1|0\.\.12" | sed "s/^/cap$v /"
done | tee gpurun_out/ac_probe.txt
timeout 300 python -m pytest tests -m gpu -q -x -k "loss or dfpn" 2>&1 | tail -2
