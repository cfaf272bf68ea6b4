#!/bin/bash
# round 2, call V: fused correlation + L1 (8f-3) parity and cfg3 step, smoke, default bench
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/v_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/v_pytest.log
python __graft_entry__.py smoke > gpurun_out/v_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/v_smoke.log
B="--steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2"
for wl in cfg3 align; do
  timeout 300 python bench.py --workload $wl $B > gpurun_out/v_$wl.json 2> gpurun_out/v_$wl.err; echo "$wl rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/v_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        ks=" ".join("%s=%.1f"%(k["call"],k["avg_us"]) for k in d.get("kernels",[]))
        print(f.split("/")[-1], "step_us %.1f"%(d["ms_per_step"]*1e3), "roofline %.3f"%d.get("roofline",{}).get("frac",0), ks)
    except Exception as e: print(f,"ERR",e)
PY
