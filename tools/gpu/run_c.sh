#!/bin/bash
# round 2, call C: parity suite after fixes, ncu --set full of the headline kernels
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_pytest.log
tail -15 gpurun_out/c_pytest.log
python tools/gpu/prof_kernels.py > gpurun_out/c_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"corr_tc_kernel|warp_fwd_kernel|warp_staged_kernel" -s 9 -c 6 \
    -o gpurun_out/c_prof -f python tools/gpu/prof_kernels.py > gpurun_out/c_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/c_ncu.log
