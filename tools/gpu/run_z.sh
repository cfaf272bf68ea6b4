#!/bin/bash
# round 2, call Z: in-context cost of every cfg3 launch (prefix timing), ticket fold vs the separate fold kernel
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/gpu/probe_cfg3_step.py > gpurun_out/z_probe_ticket.txt 2>&1
MT_B200_LIB=$PWD/master_thesis_b200/libmt_finish.so python tools/gpu/probe_cfg3_step.py > gpurun_out/z_probe_finish.txt 2>&1
MT_PDL=0 python tools/gpu/probe_cfg3_step.py > gpurun_out/z_probe_ticket_nopdl.txt 2>&1
paste gpurun_out/z_probe_ticket.txt gpurun_out/z_probe_finish.txt | cut -c1-200
cat gpurun_out/z_probe_ticket_nopdl.txt
