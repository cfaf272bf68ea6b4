#!/bin/bash
# round 2, call AM: cfg1 at 128 frames - step vs sum of the isolated calls after the PDL trigger change
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python bench.py --workload cfg1 --batch 128 --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2 > gpurun_out/am_cfg1b128.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open("gpurun_out/am_cfg1b128.json").read().strip().splitlines()[-1])
print("cfg1 B=128 step_us %.1f"%(d["ms_per_step"]*1e3), " ".join("%s=%.1f(%.2f)"%(k["call"],k["avg_us"],k["frac_hbm"]) for k in d["kernels"]))
PY
