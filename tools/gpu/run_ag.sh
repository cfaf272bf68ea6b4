#!/bin/bash
# round 2, call AG: composite fwd / bwd and l1x3 bwd with every load ahead of the first store
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "chn or l1 or inpaint" 2>&1 | tail -2
B="--steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2"
for r in 1 2; do timeout 300 python bench.py --workload cfg5 $B > gpurun_out/ag_cfg5_$r.json 2>/dev/null; done
timeout 300 python bench.py --workload cfg4 $B > gpurun_out/ag_cfg4.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/ag_*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split("/")[-1], "step_us %.1f"%(d["ms_per_step"]*1e3), " ".join("%s=%.1f"%(k["call"],k["avg_us"]) for k in d["kernels"]))
PY
