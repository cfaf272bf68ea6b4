#!/bin/bash
# round 2, call T: flow-pack parity, paired tap loads (libmt_pair.so = -DMT_TAP_PAIR=1) vs default, 64x16 staged tiles
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/t_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/t_pytest.log
MT_B200_LIB=$PWD/master_thesis_b200/libmt_pair.so timeout 600 python -m pytest tests -m gpu -q -x -k "align or warp or loss or lowres or inpaint or dfpn" > gpurun_out/t_pytest_pair.log 2>&1; echo "pytest pair rc=$?"
tail -2 gpurun_out/t_pytest_pair.log
B="--steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2"
for wl in cfg1 cfg3 align; do
  for rep in 1 2; do
  timeout 300 python bench.py --workload $wl $B > gpurun_out/t_${wl}_base$rep.json 2> gpurun_out/t_${wl}_base$rep.err
  MT_B200_LIB=$PWD/master_thesis_b200/libmt_pair.so timeout 300 python bench.py --workload $wl $B > gpurun_out/t_${wl}_pair$rep.json 2> gpurun_out/t_${wl}_pair$rep.err
  done
done
for v in "32 32 80" "64 32 80" "64 16 80" "64 16 112"; do
  set -- $v
  MT_WARP_TILE_W=$1 MT_WARP_TILE_H=$2 MT_WARP_BOX_W=$3 timeout 300 python bench.py --workload cfg2 $B > gpurun_out/t_cfg2_$1x$2_$3.json 2> gpurun_out/t_cfg2_$1x$2_$3.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/t_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        ks=" ".join("%s=%.1f"%(k["call"],k["avg_us"]) for k in d.get("kernels",[]))
        print(f.split("/")[-1], "step_us %.1f"%(d["ms_per_step"]*1e3), "roofline %.3f"%d.get("roofline",{}).get("frac",0), ks)
    except Exception as e: print(f,"ERR",e)
PY
