#!/bin/bash
# round 2, call AN (last seconds of the budget): CM groups with no parked batches
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="--steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2"
MT_CM_GROUPS=16 timeout 100 python bench.py --workload cfg2 $B > gpurun_out/an_g16.json 2>/dev/null
MT_CM_GROUPS=4 timeout 100 python bench.py --workload cfg2 $B > gpurun_out/an_g4.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/an_*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split("/")[-1], "step_us %.1f"%(d["ms_per_step"]*1e3), " ".join("%s=%.1f"%(k["call"],k["avg_us"]) for k in d["kernels"]))
PY
