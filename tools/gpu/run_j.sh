#!/bin/bash
# round 2, call J (8 GPUs): scaling of the configs that name multi-GPU runs (BASELINE configs[2..4]) + default workload at 8
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/j_topo.txt 2>&1
run() {  # name, nproc, args...
  name=$1; n=$2; shift 2
  if [ "$n" = "1" ]; then
    timeout 300 python bench.py --gpus 1 "$@" > gpurun_out/j_${name}_n$n.json 2> gpurun_out/j_${name}_n$n.err
  else
    NCCL_DEBUG=INFO timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
      bench.py --gpus $n "$@" > gpurun_out/j_${name}_n$n.json 2> gpurun_out/j_${name}_n$n.err
  fi
  echo "$name n=$n rc=$?"
}
COMMON="--steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 20"
run align 8 $COMMON
run cfg3 1 --workload cfg3 $COMMON
run cfg3 8 --workload cfg3 $COMMON
run cfg5 1 --workload cfg5 $COMMON
run cfg5 8 --workload cfg5 $COMMON
for n in 1 2 4 8; do run cfg4 $n --workload cfg4 $COMMON; done
run align 2 $COMMON
run align 4 $COMMON
grep -h -E "NVLS|Connected all|comm 0x.* rank 0 nranks" gpurun_out/j_cfg3_n8.err | head -8 > gpurun_out/j_nccl_lines.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/j_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "n", d["n_gpus"], "step_us %.1f"%(d["ms_per_step"]*1e3), "value %.0f"%d["value"], "e2e %.0f"%d["e2e"]["value"], "copy_frac %.2f"%d["e2e"].get("fraction_of_copy_ceiling",0), "gbs/rank %.1f"%d["e2e"].get("copies_only_gbs_per_rank",0), d.get("ddp",{}).get("allreduce_bytes_per_step"))
    except Exception as e: print(f,"ERR",e)
PY
