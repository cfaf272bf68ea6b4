#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/gpu/probe_cfg3_step.py subsets > gpurun_out/z2_ticket.txt 2>&1
MT_B200_LIB=$PWD/master_thesis_b200/libmt_finish.so python tools/gpu/probe_cfg3_step.py subsets > gpurun_out/z2_finish.txt 2>&1
paste gpurun_out/z2_ticket.txt gpurun_out/z2_finish.txt | cut -c1-160
