#!/bin/bash
# round 2, call AJ: does the DFPN warp's early trigger (pre-launching the 200 KB-smem staged kernel / next correlation) cost anything?
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="--steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2"
for wl in align cfg1 cfg4; do
  timeout 300 python bench.py --workload $wl $B > gpurun_out/aj_${wl}_t1.json 2>/dev/null
  MT_WARP_EARLY_TRIGGER=0 timeout 300 python bench.py --workload $wl $B > gpurun_out/aj_${wl}_t0.json 2>/dev/null
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/aj_*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split("/")[-1], "step_us %.1f"%(d["ms_per_step"]*1e3), " ".join("%s=%.1f"%(k["call"],k["avg_us"]) for k in d["kernels"]))
PY
