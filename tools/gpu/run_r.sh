#!/bin/bash
# round 2, call R: grouped CM kernel v5 (dynamic batch hand-out inside the group) - parity, bench, timeline
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "cm or guard or shard or smoke" > gpurun_out/r_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/r_pytest.log
tail -4 gpurun_out/r_pytest.log
if [ $rc -ne 0 ]; then exit 0; fi
for b in 8 32 64; do
  timeout 120 python bench.py --workload cfg2 --batch $b --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2 \
    > gpurun_out/r_cm_b$b.json 2> gpurun_out/r_cm_b$b.err
  echo "b=$b rc=$?"
done
timeout 120 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2 > gpurun_out/r_align.json 2> gpurun_out/r_align.err; echo "align rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "step_us %.1f"%(d["ms_per_step"]*1e3), [(k["call"],round(k["avg_us"],1)) for k in d["kernels"]])
    except Exception as e: print(f,"ERR",e)
PY
python tools/dbg_cm_timeline.py 8 > gpurun_out/r_cm_timeline.txt 2>&1; tail -4 gpurun_out/r_cm_timeline.txt
