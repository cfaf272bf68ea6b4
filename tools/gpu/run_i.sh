#!/bin/bash
# round 2, call I: full GPU suite + default bench (new corr heuristic, NUMA binding, copy ceiling) + cfg5
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/i_pytest.log
tail -4 gpurun_out/i_pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/i_align.json 2> gpurun_out/i_align.err; echo "align rc=$?"
python bench.py --workload cfg5 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/i_cfg5.json 2> gpurun_out/i_cfg5.err; echo "cfg5 rc=$?"
python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/i_cfg2.json 2> gpurun_out/i_cfg2.err; echo "cfg2 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/i_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "step_us %.1f"%(d["ms_per_step"]*1e3), "value %.0f"%d["value"], [(k["call"],round(k["avg_us"],1)) for k in d["kernels"]])
        print("   e2e", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["e2e"].items() if k!="api"}, d.get("host_affinity"))
    except Exception as e: print(f,"ERR",e)
PY
nvidia-smi topo -m > gpurun_out/i_topo.txt 2>&1
