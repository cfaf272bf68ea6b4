#!/bin/bash
# round 2, call M: CTA-pair correlation with a relaxed relay - parity + sweep
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "cta_pairs" > gpurun_out/m_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/m_pytest.log
tail -3 gpurun_out/m_pytest.log
for b in 8 32 128; do
  for t in "0 0" "1 256" "1 128"; do
    set -- $t
    MT_CORR_2CTA=$1 MT_CORR_TN=$2 timeout 120 python bench.py --workload cfg1 --batch $b --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2 \
      > gpurun_out/m_corr_b${b}_pair$1_tn$2.json 2> gpurun_out/m_corr_b${b}_pair$1_tn$2.err
    echo "b=$b pair=$1 tn=$2 rc=$?"
  done
done
for t in "0 0" "1 256" "1 128"; do
  set -- $t
  MT_CORR_2CTA=$1 MT_CORR_TN=$2 timeout 120 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2 \
      > gpurun_out/m_align_pair$1_tn$2.json 2> gpurun_out/m_align_pair$1_tn$2.err
  echo "align pair=$1 tn=$2 rc=$?"
  MT_CORR_2CTA=$1 MT_CORR_TN=$2 timeout 120 python bench.py --workload cfg3 --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 2 \
      > gpurun_out/m_cfg3_pair$1_tn$2.json 2> gpurun_out/m_cfg3_pair$1_tn$2.err
  echo "cfg3 pair=$1 tn=$2 rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/m_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "step_us %.1f"%(d["ms_per_step"]*1e3), [(k["call"],round(k["avg_us"],1)) for k in d["kernels"] if "corr" in k["call"]])
    except Exception as e: print(f,"ERR",e)
PY
MT_CORR_2CTA=1 python tools/gpu/prof_kernels.py corr > gpurun_out/m_plain.log 2>&1 && \
MT_CORR_2CTA=1 ncu --set full --clock-control none --import-source on -k regex:"corr_tc2" -s 3 -c 1 \
    -o gpurun_out/m_prof -f python tools/gpu/prof_kernels.py corr > gpurun_out/m_ncu.log 2>&1
echo "ncu rc=$?"
