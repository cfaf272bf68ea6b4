#!/bin/bash
# round 2, call AK: CM knobs re-checked under the new PDL behaviour (cfg2 = staged warp + CM)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="--steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2"
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --workload cfg2 $B > gpurun_out/ak_$tag.json 2>/dev/null; }
run base X=1
run g4 MT_CM_GROUPS=4
run g16 MT_CM_GROUPS=16
run keep0 MT_CM_KEEP=0
run keep2 MT_CM_KEEP=2
run stg1 MT_STAGED_EARLY_TRIGGER=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/ak_*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split("/")[-1], "step_us %.1f"%(d["ms_per_step"]*1e3), " ".join("%s=%.1f"%(k["call"],k["avg_us"]) for k in d["kernels"]))
PY
