#!/bin/bash
# 2-GPU sanity run of both bench arms as the driver launches them
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/n2_align.json 2> gpurun_out/n2_align.err; echo "align rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --workload cfg3 > gpurun_out/n2_cfg3.json 2> gpurun_out/n2_cfg3.err; echo "cfg3 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 5 --warmup 1 > gpurun_out/n2_ref.json 2> gpurun_out/n2_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
for f in ("n2_align","n2_cfg3","n2_ref"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, d.get("n_gpus"), "value %.0f"%d["value"], "step_us %.1f"%(d["ms_per_step"]*1e3), "e2e %.0f"%d["e2e"]["value"], d.get("ddp",{}) and {k:d["ddp"][k] for k in list(d["ddp"])[:4]})
    except Exception as e: print(f,"ERR",e)
PY
