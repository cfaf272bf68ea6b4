#!/bin/bash
# round 2, call AL: CM without parked batches as the default: parity, default workload, cfg2 at B = 8 / 32
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "cm or guard or bounds" 2>&1 | tail -2
B="--steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2"
timeout 300 python bench.py $B > gpurun_out/al_align.json 2>/dev/null
timeout 300 python bench.py --workload cfg2 $B > gpurun_out/al_cfg2.json 2>/dev/null
timeout 300 python bench.py --workload cfg2 --batch 32 $B > gpurun_out/al_cfg2b32_k0.json 2>/dev/null
MT_CM_KEEP=2 timeout 300 python bench.py --workload cfg2 --batch 32 $B > gpurun_out/al_cfg2b32_k2.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/al_*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split("/")[-1], "step_us %.1f"%(d["ms_per_step"]*1e3), "value %.0f"%d["value"], " ".join("%s=%.1f"%(k["call"],k["avg_us"]) for k in d["kernels"]))
PY
