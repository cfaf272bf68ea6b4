#!/bin/bash
# round 2, call E: ncu --set full of the headline kernels (corr 128 frames, DFPN warp, staged CPN warp, CM grouped)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/gpu/prof_kernels.py > gpurun_out/e_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"corr_tc_kernel|warp_fwd_kernel|warp_staged_kernel|cm_group_kernel|cm_masks" -s 15 -c 5 \
    -o gpurun_out/e_prof -f python tools/gpu/prof_kernels.py > gpurun_out/e_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/e_ncu.log
