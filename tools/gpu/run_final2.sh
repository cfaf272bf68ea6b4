#!/bin/bash
# round 2, final (second session): full GPU suite, smoke, both bench arms, per-config lines, launch list of the default
# bench, ncu --set full of the changed kernels (DFPN warp with tap reuse, L1-mode correlation + backward, flow pack)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/g2_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/g2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g2_pytest.log
tail -4 gpurun_out/g2_pytest.log
python __graft_entry__.py smoke > gpurun_out/g2_smoke.log 2>&1; echo "smoke rc=$?"
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > gpurun_out/g2_bench_reference.json 2> gpurun_out/g2_bench_reference.err; echo "reference rc=$?"
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/g2_bench_align.json 2> gpurun_out/g2_bench_align.err; echo "align rc=$?"
for wl in cfg1 cfg2 cfg3 cfg4 cfg5; do
  timeout 300 python bench.py --workload $wl --steps 50 --warmup 5 > gpurun_out/g2_bench_$wl.json 2> gpurun_out/g2_bench_$wl.err; echo "$wl rc=$?"
done
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 2 --no-graph > gpurun_out/g2_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/g2_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 2 --no-graph > gpurun_out/g2_ncu.log 2>&1
echo "ncu launches rc=$?"
python tools/gpu/prof_kernels.py dfpn corrl1 flowpack > gpurun_out/g2_prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"corr_tc|corr_l1|warp_fwd_kernel|flow_pack" -s 12 -c 6 \
    -o gpurun_out/g2_prof -f python tools/gpu/prof_kernels.py dfpn corrl1 flowpack > gpurun_out/g2_prof_ncu.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/g2_prof_ncu.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/g2_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.0f"%d["value"], "step_us %.1f"%(d["ms_per_step"]*1e3), "e2e %.0f"%d["e2e"]["value"], "roofline %.3f"%d.get("roofline",{}).get("frac",0), "cpu", round(d.get("cpu_baseline",{}).get("value",0)))
    except Exception as e: print(f,"ERR",e)
PY
