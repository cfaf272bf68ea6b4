#!/bin/bash
# round 2, call L: CTA-pair correlation kernel - try_wait vs test_wait for the cross-CTA barriers, ncu of the pair kernel
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for lib in libmt_b200.so libmt_b200_testwait.so; do
  for b in 32 128; do
    MT_B200_LIB=$PWD/master_thesis_b200/$lib MT_CORR_2CTA=1 MT_CORR_TN=256 timeout 120 python bench.py --workload cfg1 --batch $b --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2 \
      > gpurun_out/l_${lib%.so}_b$b.json 2> gpurun_out/l_${lib%.so}_b$b.err
    echo "$lib b=$b rc=$?"
  done
  MT_B200_LIB=$PWD/master_thesis_b200/$lib MT_CORR_2CTA=1 MT_CORR_TN=256 timeout 120 python bench.py --workload cfg3 --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 2 \
      > gpurun_out/l_${lib%.so}_cfg3.json 2> gpurun_out/l_${lib%.so}_cfg3.err
  echo "$lib cfg3 rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/l_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "step_us %.1f"%(d["ms_per_step"]*1e3), [(k["call"],round(k["avg_us"],1)) for k in d["kernels"] if "corr" in k["call"]])
    except Exception as e: print(f,"ERR",e)
PY
MT_CORR_2CTA=1 python tools/gpu/prof_kernels.py corr > gpurun_out/l_plain.log 2>&1 && \
MT_CORR_2CTA=1 ncu --set full --clock-control none --import-source on -k regex:"corr_tc2" -s 3 -c 1 \
    -o gpurun_out/l_prof -f python tools/gpu/prof_kernels.py corr > gpurun_out/l_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/l_ncu.log
