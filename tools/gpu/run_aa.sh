#!/bin/bash
# round 2, call AA: masked_l1 with folded axes, NULL = all-ones mask, 4 chunks in flight per thread
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "l1 or loss or dfpn or chn or inpaint" > gpurun_out/aa_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/aa_pytest.log
B="--steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2"
for wl in cfg3 cfg5; do timeout 300 python bench.py --workload $wl $B > gpurun_out/aa_$wl.json 2> gpurun_out/aa_$wl.err; done
python tools/gpu/probe_cfg3_step.py > gpurun_out/aa_probe.txt 2>&1; cat gpurun_out/aa_probe.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/aa_*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split("/")[-1], "step_us %.1f"%(d["ms_per_step"]*1e3), " ".join("%s=%.1f(%.2f)"%(k["call"],k["avg_us"],k["frac_hbm"]) for k in d["kernels"]))
PY
