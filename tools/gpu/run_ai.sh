#!/bin/bash
# round 2, call AI: new correlation tile heuristic; knobs whose optimum may have moved with the PDL trigger change
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "corr" 2>&1 | tail -2
B="--steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2"
run() { tag=$1; wl=$2; shift 2; env "$@" timeout 300 python bench.py --workload $wl $B > gpurun_out/ai_${wl}_$tag.json 2>/dev/null; }
run base align X=1
run iters2 align MT_WARP_ITERS=2
run base cfg1 X=1
run t128x128 cfg1 MT_CORR_TM=128 MT_CORR_TN=128
run iters2 cfg1 MT_WARP_ITERS=2
timeout 300 python bench.py --workload cfg1 --batch 16 $B > gpurun_out/ai_cfg1b16_base.json 2>/dev/null
MT_CORR_TM=256 MT_CORR_TN=128 timeout 300 python bench.py --workload cfg1 --batch 16 $B > gpurun_out/ai_cfg1b16_t256x128.json 2>/dev/null
MT_CORR_TM=128 MT_CORR_TN=64 timeout 300 python bench.py --workload cfg1 --batch 16 $B > gpurun_out/ai_cfg1b16_t128x64.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/ai_*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split("/")[-1], "step_us %.1f"%(d["ms_per_step"]*1e3), " ".join("%s=%.1f"%(k["call"],k["avg_us"]) for k in d["kernels"]))
PY
