set -x
CMD="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:warp_staged -s 5 -c 2 -f -o gpurun_out/prof_staged_b8 $CMD > gpurun_out/ncu1.log 2>&1
$CMD --batch 32 > gpurun_out/plain32.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:warp_staged -s 5 -c 1 -f -o gpurun_out/prof_staged_b32 $CMD --batch 32 > gpurun_out/ncu2.log 2>&1
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/launches_cfg2.csv $CMD > gpurun_out/ncu3.log 2>&1
tail -3 gpurun_out/ncu1.log gpurun_out/ncu2.log gpurun_out/ncu3.log
