#!/bin/bash
# round 2, call AD: loss-only fused forward at 64 registers (8 CTAs / SM), CTA cap 6 / 8 / 16
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in 6 8 16; do
  MT_WARPL1_CTAS_PER_SM=$v python tools/gpu/probe_cfg3_step.py 2>&1 | grep -E "warp_l1_fwd|0\.\.12" | sed "s/^/cap$v /"
done | tee gpurun_out/ad_probe.txt
timeout 300 python -m pytest tests -m gpu -q -x -k "loss or dfpn or align" 2>&1 | tail -2
