#!/bin/bash
# round 2, call B: full GPU parity suite (no -x), corr tile sweep at 8 / 40 / 128 frames
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
tail -40 gpurun_out/b_pytest.log
for tn in 0 64 128 256; do
  for wl in "cfg1 8" "cfg1 40" "cfg3 32"; do
    set -- $wl
    MT_CORR_TN=$tn timeout 300 python bench.py --workload $1 --batch $2 --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 2 \
      > gpurun_out/b_corr_tn${tn}_$1_$2.json 2> gpurun_out/b_corr_tn${tn}_$1_$2.err
    echo "tn=$tn $1 b=$2 rc=$?"
  done
done
