#!/bin/bash
# round 2, call AE: staged CPN warp with two CTAs per SM (2-stage rings) vs one CTA with 5 stages; 2-GPU sanity run
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "staged" 2>&1 | tail -2
B="--steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2"
for wl in cfg2 cfg5 align; do
  timeout 300 python bench.py --workload $wl $B > gpurun_out/ae_${wl}_c1.json 2>/dev/null
  MT_WARP_STAGED_CTAS=2 timeout 300 python bench.py --workload $wl $B > gpurun_out/ae_${wl}_c2.json 2>/dev/null
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/ae_*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split("/")[-1], "step_us %.1f"%(d["ms_per_step"]*1e3), " ".join("%s=%.1f"%(k["call"],k["avg_us"]) for k in d["kernels"]))
PY
