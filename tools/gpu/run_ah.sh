#!/bin/bash
# round 2, call AH: correlation tile shape at 32 frames inside the default workload, after the PDL trigger change
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="--steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2"
run() { tag=$1; shift; env "$@" timeout 300 python bench.py $B > gpurun_out/ah_$tag.json 2>/dev/null; }
run base X=1
run t128x128 MT_CORR_TM=128 MT_CORR_TN=128
run t128x64 MT_CORR_TM=128 MT_CORR_TN=64
run t256x64 MT_CORR_TM=256 MT_CORR_TN=64
run t128x256 MT_CORR_TM=128 MT_CORR_TN=256
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/ah_*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split("/")[-1], "step_us %.1f"%(d["ms_per_step"]*1e3), " ".join("%s=%.1f"%(k["call"],k["avg_us"]) for k in d["kernels"]))
PY
