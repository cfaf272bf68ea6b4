#!/bin/bash
# round 2, call AF: chn_l1x3_fwd with a register budget (all loads in flight) and a one-wave grid
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "chn or l1" 2>&1 | tail -2
B="--steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2"
for c in 3 2 4 6 8; do
  MT_L1X3_CTAS_PER_SM=$c timeout 300 python bench.py --workload cfg5 $B > gpurun_out/af_cfg5_c$c.json 2>/dev/null
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/af_*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split("/")[-1], "step_us %.1f"%(d["ms_per_step"]*1e3), " ".join("%s=%.1f"%(k["call"],k["avg_us"]) for k in d["kernels"]))
PY
