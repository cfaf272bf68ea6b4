#!/bin/bash
# round 2, call D: grouped CM kernel - parity, then cfg2 bench per launch structure / group count
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "cm or guard" > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest.log
tail -5 gpurun_out/d_pytest.log
for cfg in "1 0 8" "2 0 8" "2 4 8" "2 2 8" "1 0 32" "2 0 32" "2 4 32" "2 16 32"; do
  set -- $cfg
  MT_CM_TABLE=$1 MT_CM_GROUPS=$2 timeout 300 python bench.py --workload cfg2 --batch $3 --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2 \
    > gpurun_out/d_cm_t$1_g$2_b$3.json 2> gpurun_out/d_cm_t$1_g$2_b$3.err
  echo "table=$1 groups=$2 b=$3 rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/d_cm_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "step_us %.1f"%(d["ms_per_step"]*1e3), [(k["call"],round(k["avg_us"],1)) for k in d["kernels"]])
    except Exception as e: print(f,"ERR",e)
PY
