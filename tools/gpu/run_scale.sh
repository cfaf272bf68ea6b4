#!/bin/bash
# round 2: scaling runs on an N-GPU box (N = $1): the configs that name multi-GPU runs (BASELINE configs[2..4]) + the
# default workload.  Every run has its own short timeout.
cd $GRAFT_REPO_ROOT
NMAX=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/s${NMAX}_topo.txt 2>&1
run() {  # name, nproc, args...
  name=$1; n=$2; shift 2
  if [ "$n" = "1" ]; then
    timeout 150 python bench.py --gpus 1 "$@" > gpurun_out/s${NMAX}_${name}_n$n.json 2> gpurun_out/s${NMAX}_${name}_n$n.err
  else
    NCCL_DEBUG=INFO timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
      bench.py --gpus $n "$@" > gpurun_out/s${NMAX}_${name}_n$n.json 2> gpurun_out/s${NMAX}_${name}_n$n.err
  fi
  echo "$name n=$n rc=$?"
}
COMMON="--steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 20"
NS="1"; for n in 2 4 8; do [ $n -le $NMAX ] && NS="$NS $n"; done
for n in $NS; do run cfg3 $n --workload cfg3 $COMMON; done
for n in $NS; do run cfg5 $n --workload cfg5 $COMMON; done
# the same without the gradient all-reduce (the hot-path kernels of a step last 0.1 - 0.4 ms, a 56 MB all-reduce as long or longer)
run cfg3noddp $NMAX --workload cfg3 --ddp-mb 0 $COMMON
run cfg5noddp $NMAX --workload cfg5 --ddp-mb 0 $COMMON
for n in $NS; do run cfg4 $n --workload cfg4 $COMMON; done
for n in $NS; do run align $n $COMMON; done
grep -h -E "NVLS|Connected all|comm 0x.* rank 0 nranks|via NVL" gpurun_out/s${NMAX}_cfg3_n${NMAX}.err | head -8 > gpurun_out/s${NMAX}_nccl_lines.txt
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/s${NMAX}_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "n", d["n_gpus"], "step_us %.1f"%(d["ms_per_step"]*1e3), "value %.0f"%d["value"], "e2e %.0f"%d["e2e"]["value"], "copy_frac %.2f"%d["e2e"].get("fraction_of_copy_ceiling",0), "gbs/rank %.1f"%d["e2e"].get("copies_only_gbs_per_rank",0), d.get("ddp",{}).get("allreduce_bytes_per_step"), d.get("host_affinity"))
    except Exception as e: print(f,"ERR",e)
PY
