#!/bin/bash
# weak scaling of the default workload on one box: N = 1, 2, 4, 8 (needs gpurun --gpus 8)
O=gpurun_out/scaling.jsonl; : > $O
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then python bench.py --gpus 1 --no-cpu-baseline >> $O 2>gpurun_out/scaling_err_$N.log
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --no-cpu-baseline >> $O 2>gpurun_out/scaling_err_$N.log; fi
done
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/scaling.jsonl') if l.strip().startswith('{')]
base=rows[0]['value']
for r in rows: print('N=%d value=%.0f frames/s  %.2fx  step=%.1f us  e2e=%.0f' % (r['n_gpus'], r['value'], r['value']/base, r['ms_per_step']*1e3, r['e2e']['value']))
PY
