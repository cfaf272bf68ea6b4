#!/bin/bash
# round 2, call U: vertical tap reuse in the dense-flow gather kernels (default lib) vs -DMT_TAP_REUSE=0, 2 vs 4 rows per thread
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/u_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/u_pytest.log
MT_B200_LIB=$PWD/master_thesis_b200/libmt_rows4.so timeout 600 python -m pytest tests -m gpu -q -x -k "align or warp or loss or lowres or inpaint or dfpn" > gpurun_out/u_pytest_rows4.log 2>&1; echo "pytest rows4 rc=$?"
tail -2 gpurun_out/u_pytest_rows4.log
B="--steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2"
run() { # tag, lib, workload, env...
  tag=$1; lib=$2; wl=$3; shift 3
  env "$@" MT_B200_LIB=$PWD/master_thesis_b200/$lib timeout 300 python bench.py --workload $wl $B > gpurun_out/u_${wl}_$tag.json 2> gpurun_out/u_${wl}_$tag.err
}
for wl in cfg1 align cfg4; do
  run noreuse libmt_noreuse.so $wl X=1
  run reuse2 libmt_b200.so $wl X=1
  run reuse4 libmt_b200.so $wl MT_WARP_ROWS=4
  run noreuse4 libmt_noreuse.so $wl MT_WARP_ROWS=4
done
run noreuse libmt_noreuse.so cfg3 X=1
run reuse2 libmt_b200.so cfg3 X=1
run reuse4 libmt_rows4.so cfg3 X=1
run reuse2 libmt_b200.so cfg5 X=1
run noreuse libmt_noreuse.so cfg5 X=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/u_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        ks=" ".join("%s=%.1f"%(k["call"],k["avg_us"]) for k in d.get("kernels",[]))
        print(f.split("/")[-1], "step_us %.1f"%(d["ms_per_step"]*1e3), "roofline %.3f"%d.get("roofline",{}).get("frac",0), ks)
    except Exception as e: print(f,"ERR",e)
PY
