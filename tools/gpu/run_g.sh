#!/bin/bash
# round 2, call G: grouped CM kernel v3 + corr_tc with LDS (shared address space kept): parity, benches, ncu
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g_pytest.log
tail -4 gpurun_out/g_pytest.log
for cfg in "1 8" "2 8" "1 32" "2 32"; do
  set -- $cfg
  MT_CM_TABLE=$1 timeout 300 python bench.py --workload cfg2 --batch $2 --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2 \
    > gpurun_out/g_cm_t$1_b$2.json 2> gpurun_out/g_cm_t$1_b$2.err
  echo "table=$1 b=$2 rc=$?"
done
python bench.py --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2 > gpurun_out/g_align.json 2> gpurun_out/g_align.err; echo "align rc=$?"
python bench.py --workload cfg3 --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/g_cfg3.json 2> gpurun_out/g_cfg3.err; echo "cfg3 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/g_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "step_us %.1f"%(d["ms_per_step"]*1e3), [(k["call"],round(k["avg_us"],1)) for k in d["kernels"]])
    except Exception as e: print(f,"ERR",e)
PY
python tools/gpu/prof_kernels.py cm corr > gpurun_out/g_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"cm_group_kernel|corr_tc" -s 6 -c 2 \
    -o gpurun_out/g_prof -f python tools/gpu/prof_kernels.py cm corr > gpurun_out/g_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/g_ncu.log
