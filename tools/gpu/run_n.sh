#!/bin/bash
# round 2, call N: CM grouped kernel v4 (pipelined across samples) - parity, A/B bench, sanitizer on small cases
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "cm or guard or shard or smoke" > gpurun_out/n_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/n_pytest.log
tail -4 gpurun_out/n_pytest.log
if [ $rc -ne 0 ]; then exit 0; fi
for cfg in "0 8" "-1 8" "0 32" "-1 32" "-1 64" "0 64"; do
  set -- $cfg
  MT_CM_PIPE=$1 timeout 120 python bench.py --workload cfg2 --batch $2 --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2 \
    > gpurun_out/n_cm_pipe$1_b$2.json 2> gpurun_out/n_cm_pipe$1_b$2.err
  echo "pipe=$1 b=$2 rc=$?"
done
timeout 120 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2 > gpurun_out/n_align.json 2> gpurun_out/n_align.err; echo "align rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/n_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "step_us %.1f"%(d["ms_per_step"]*1e3), [(k["call"],round(k["avg_us"],1)) for k in d["kernels"]])
    except Exception as e: print(f,"ERR",e)
PY
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/gpu/sanitize_small.py > gpurun_out/n_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/n_memcheck.log
