#!/bin/bash
# round 2, call Z3: smem-heavy persistent kernels (corr, staged warp) that let their dependents pre-launch keep the SM in
# the max-shared-memory configuration for the gather kernels that follow: early trigger on / off
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="--steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2"
run() { tag=$1; wl=$2; shift 2; env "$@" timeout 300 python bench.py --workload $wl $B > gpurun_out/z3_${wl}_$tag.json 2> gpurun_out/z3_${wl}_$tag.err; }
for wl in align cfg1 cfg3 cfg2 cfg5; do
  run e1 $wl X=1
  run e0 $wl MT_CORR_EARLY_TRIGGER=0
  run e0s0 $wl MT_CORR_EARLY_TRIGGER=0 MT_STAGED_EARLY_TRIGGER=0
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/z3_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "step_us %.1f"%(d["ms_per_step"]*1e3), "instr_us %.1f"%(d["rounds_ms"]["instrumented"]*1e3/d["steps"]), "value %.0f" % d["value"])
    except Exception as e: print(f,"ERR",e)
PY
MT_CORR_EARLY_TRIGGER=0 python tools/gpu/probe_cfg3_step.py > gpurun_out/z3_probe_e0.txt 2>&1; cat gpurun_out/z3_probe_e0.txt
