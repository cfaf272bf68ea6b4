#!/bin/bash
# round 2 final: full GPU suite, smoke, driver-style bench (both arms), per-config lines, launch list of the default bench
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/f2_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/f2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f2_pytest.log
tail -4 gpurun_out/f2_pytest.log
python __graft_entry__.py smoke > gpurun_out/f2_smoke.log 2>&1; echo "smoke rc=$?"
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > gpurun_out/f2_bench_reference.json 2> gpurun_out/f2_bench_reference.err; echo "reference rc=$?"
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/f2_bench_align.json 2> gpurun_out/f2_bench_align.err; echo "align rc=$?"
for wl in cfg1 cfg2 cfg3 cfg4 cfg5; do
  timeout 300 python bench.py --workload $wl --steps 50 --warmup 5 > gpurun_out/f2_bench_$wl.json 2> gpurun_out/f2_bench_$wl.err; echo "$wl rc=$?"
done
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 2 --no-graph > gpurun_out/f2_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/f2_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 2 --no-graph > gpurun_out/f2_ncu.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/f2_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.0f"%d["value"], "step_us %.1f"%(d["ms_per_step"]*1e3), "e2e %.0f"%d["e2e"]["value"], "roofline %.3f"%d.get("roofline",{}).get("frac",0), "cpu", round(d.get("cpu_baseline",{}).get("value",0)))
    except Exception as e: print(f,"ERR",e)
PY
