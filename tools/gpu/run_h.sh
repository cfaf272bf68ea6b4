#!/bin/bash
# round 2, call H: corr TM x TN tile sweep (8 / 32 / 128 frames), DFPN warp row-group sweep with the next-flow prefetch
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "corr or warp or align or dfpn or smoke or guard or lowres or inference or inpaint" > gpurun_out/h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h_pytest.log
tail -4 gpurun_out/h_pytest.log
for b in 8 32 128; do
  for t in "0 0" "256 256" "256 128" "256 64" "128 256" "128 128" "128 64"; do
    set -- $t
    MT_CORR_TM=$1 MT_CORR_TN=$2 timeout 300 python bench.py --workload cfg1 --batch $b --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2 \
      > gpurun_out/h_corr_b${b}_tm$1_tn$2.json 2> gpurun_out/h_corr_b${b}_tm$1_tn$2.err
    echo "b=$b tm=$1 tn=$2 rc=$?"
  done
done
for it in 1 2 4 8; do
  for b in 8 32 128; do
    MT_WARP_ITERS=$it timeout 300 python bench.py --workload cfg1 --batch $b --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2 \
      > gpurun_out/h_warp_b${b}_it$it.json 2> gpurun_out/h_warp_b${b}_it$it.err
    echo "iters=$it b=$b rc=$?"
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/h_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "step_us %.1f"%(d["ms_per_step"]*1e3), [(k["call"],round(k["avg_us"],1)) for k in d["kernels"]])
    except Exception as e: print(f,"ERR",e)
PY
