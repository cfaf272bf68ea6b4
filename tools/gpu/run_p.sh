#!/bin/bash
# round 2, call P: fused-loss forward on consecutive row strips (L1 locality) - parity + cfg3 bench + ncu
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "loss or dfpn_training or smoke or guard or cfg3" > gpurun_out/p_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/p_pytest.log
tail -3 gpurun_out/p_pytest.log
for c in 16 32 64; do
  MT_WARPL1_CTAS_PER_SM=$c timeout 120 python bench.py --workload cfg3 --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 2 \
    > gpurun_out/p_cfg3_cap$c.json 2> gpurun_out/p_cfg3_cap$c.err
  echo "cap=$c rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/p_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "step_us %.1f"%(d["ms_per_step"]*1e3), [(k["call"],round(k["avg_us"],1)) for k in d["kernels"] if "warp_l1" in k["call"]])
    except Exception as e: print(f,"ERR",e)
PY
python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 2 --no-graph > gpurun_out/p_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"warp_l1_fwd_kernel" -s 7 -c 1 \
    -o gpurun_out/p_prof -f python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 2 --no-graph > gpurun_out/p_ncu.log 2>&1
echo "ncu rc=$?"
