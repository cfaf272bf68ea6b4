#!/bin/bash
# round 2, call W: persistent pipelined dense-flow forward (MT_WARP_PIPE) - parity and cfg1 / align A/B; cfg3 with and without PDL
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "persistent or full_size_both" > gpurun_out/w_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/w_pytest.log
B="--steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2"
run() { tag=$1; wl=$2; shift 2; env "$@" timeout 300 python bench.py --workload $wl $B > gpurun_out/w_${wl}_$tag.json 2> gpurun_out/w_${wl}_$tag.err; }
for wl in cfg1 align; do
  run base $wl MT_WARP_PIPE=0
  run pipe_r2c4 $wl MT_WARP_PIPE=1
  run pipe_r2c3 $wl MT_WARP_PIPE=1 MT_WARP_PIPE_CTAS=3
  run pipe_r2c2 $wl MT_WARP_PIPE=1 MT_WARP_PIPE_CTAS=2
  run pipe_r1c4 $wl MT_WARP_PIPE=1 MT_WARP_PIPE_ROWS=1
  run pipe_r1c5 $wl MT_WARP_PIPE=1 MT_WARP_PIPE_ROWS=1 MT_WARP_PIPE_CTAS=5
done
run b32_base cfg1 MT_WARP_PIPE=0 X=1 -- 2>/dev/null
timeout 300 python bench.py --workload cfg1 --batch 32 $B > gpurun_out/w_cfg1b32_base.json 2>/dev/null
MT_WARP_PIPE=1 timeout 300 python bench.py --workload cfg1 --batch 32 $B > gpurun_out/w_cfg1b32_pipe.json 2>/dev/null
run pdl1 cfg3 X=1
run pdl0 cfg3 MT_PDL=0
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/w_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        ks=" ".join("%s=%.1f"%(k["call"],k["avg_us"]) for k in d.get("kernels",[]))
        print(f.split("/")[-1], "step_us %.1f"%(d["ms_per_step"]*1e3), "instr_us %.1f"%(d["rounds_ms"]["instrumented"]*1e3/d["steps"]), "roofline %.3f"%d.get("roofline",{}).get("frac",0), ks)
    except Exception as e: print(f,"ERR",e)
PY
