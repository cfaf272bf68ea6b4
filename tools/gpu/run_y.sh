#!/bin/bash
# round 2, call Y: L1-mode correlation with the in-kernel fold (one launch), coalesced backward; tile variants at 128 / 32 frames
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "corr or dfpn_loss or compute_loss or train_val" > gpurun_out/y_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/y_pytest.log
for b in 32 8; do
  python tools/gpu/time_corr_l1.py $b
  MT_CORR_2CTA=1 python tools/gpu/time_corr_l1.py $b
  MT_CORR_TM=256 MT_CORR_TN=128 python tools/gpu/time_corr_l1.py $b
  MT_CORR_TM=128 MT_CORR_TN=128 python tools/gpu/time_corr_l1.py $b
  MT_CORR_TM=128 MT_CORR_TN=256 python tools/gpu/time_corr_l1.py $b
done 2>&1 | grep frames | tee gpurun_out/y_corr_l1_sweep.txt
timeout 300 python bench.py --workload cfg3 --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2 > gpurun_out/y_cfg3.json 2> gpurun_out/y_cfg3.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/y_cfg3.json").read().strip().splitlines()[-1])
print("cfg3 step_us %.1f"%(d["ms_per_step"]*1e3), " ".join("%s=%.1f"%(k["call"],k["avg_us"]) for k in d["kernels"]))
PY
