#!/bin/bash
# round 2, call F: grouped CM kernel v2 (c_t in pass 1, on-chip retention) - parity, bench sweep, ncu
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "cm or guard or smoke" > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log
tail -5 gpurun_out/f_pytest.log
for cfg in "1 8 8" "2 8 8" "2 0 8" "2 1 8" "1 8 32" "2 8 32" "2 8 64"; do
  set -- $cfg
  MT_CM_TABLE=$1 MT_CM_KEEP=$2 timeout 300 python bench.py --workload cfg2 --batch $3 --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2 \
    > gpurun_out/f_cm_t$1_k$2_b$3.json 2> gpurun_out/f_cm_t$1_k$2_b$3.err
  echo "table=$1 keep=$2 b=$3 rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/f_cm_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "step_us %.1f"%(d["ms_per_step"]*1e3), [(k["call"],round(k["avg_us"],1)) for k in d["kernels"]])
    except Exception as e: print(f,"ERR",e)
PY
python tools/gpu/prof_kernels.py cm > gpurun_out/f_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"cm_group_kernel|cm_masks" -s 6 -c 2 \
    -o gpurun_out/f_prof -f python tools/gpu/prof_kernels.py cm > gpurun_out/f_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/f_ncu.log
