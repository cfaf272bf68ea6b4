#!/bin/bash
# round 2, call A: GPU parity suite, smoke, bench (default + cfg3), launch list of the default bench
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -25 gpurun_out/a_pytest.log
python __graft_entry__.py smoke > gpurun_out/a_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 20 --warmup 3 > gpurun_out/a_bench_align.json 2> gpurun_out/a_bench_align.err; echo "bench rc=$?"
python bench.py --workload cfg3 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_cfg3.json 2> gpurun_out/a_bench_cfg3.err; echo "cfg3 rc=$?"
python bench.py --workload cfg4 --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_cfg4.json 2> gpurun_out/a_bench_cfg4.err; echo "cfg4 rc=$?"
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/a_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/a_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/a_ncu.log 2>&1
echo "ncu rc=$?"
