#!/bin/bash
# round 2, call O: cfg3 fused-loss forward CTA-cap sweep + ncu of the 256x256 warp_l1 fwd / bwd launches
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for c in 8 16 32 64 110; do
  MT_WARPL1_CTAS_PER_SM=$c timeout 120 python bench.py --workload cfg3 --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 2 \
    > gpurun_out/o_cfg3_cap$c.json 2> gpurun_out/o_cfg3_cap$c.err
  echo "cap=$c rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/o_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "step_us %.1f"%(d["ms_per_step"]*1e3), [(k["call"],round(k["avg_us"],1)) for k in d["kernels"] if "warp_l1" in k["call"]])
    except Exception as e: print(f,"ERR",e)
PY
python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 2 --no-graph > gpurun_out/o_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"warp_l1_(fwd|bwd)_kernel" -s 12 -c 4 \
    -o gpurun_out/o_prof -f python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 2 --no-graph > gpurun_out/o_ncu.log 2>&1
echo "ncu rc=$?"
