#!/bin/bash
# round 2, call X: more resident warps for the dense-flow forward: 1 row per thread (48 registers, 10 CTAs / SM), MINB 7 / 8
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "rows_per_thread or full_size" > gpurun_out/x_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/x_pytest.log
B="--steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2"
run() { tag=$1; wl=$2; shift 2; env "$@" timeout 300 python bench.py --workload $wl $B > gpurun_out/x_${wl}_$tag.json 2> gpurun_out/x_${wl}_$tag.err; }
L=$PWD/master_thesis_b200
for wl in cfg1 align; do
  run base $wl X=1
  run rows1 $wl MT_WARP_ROWS=1
  run rows1i2 $wl MT_WARP_ROWS=1 MT_WARP_ITERS=2
  run minb7 $wl MT_B200_LIB=$L/libmt_minb7.so
  run minb8 $wl MT_B200_LIB=$L/libmt_minb8.so
done
timeout 300 python bench.py --workload cfg1 --batch 32 $B > gpurun_out/x_cfg1b32_base.json 2>/dev/null
MT_WARP_ROWS=1 timeout 300 python bench.py --workload cfg1 --batch 32 $B > gpurun_out/x_cfg1b32_rows1.json 2>/dev/null
MT_B200_LIB=$L/libmt_minb8.so timeout 300 python bench.py --workload cfg1 --batch 32 $B > gpurun_out/x_cfg1b32_minb8.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/x_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        ks=" ".join("%s=%.1f"%(k["call"],k["avg_us"]) for k in d.get("kernels",[]))
        print(f.split("/")[-1], "step_us %.1f"%(d["ms_per_step"]*1e3), "roofline %.3f"%d.get("roofline",{}).get("frac",0), ks)
    except Exception as e: print(f,"ERR",e)
PY
