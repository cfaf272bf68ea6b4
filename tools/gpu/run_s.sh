#!/bin/bash
# round 2, call S: tensor-pipe-active of the correlation kernels at 128 / 256 / 512 frames
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/gpu/prof_corr_scaling.py > gpurun_out/s_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size \
    --clock-control none -k regex:"corr_tc" --csv --log-file gpurun_out/s_corr_scaling.csv python tools/gpu/prof_corr_scaling.py > gpurun_out/s_ncu.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/s_corr_scaling.csv")) if len(r)>5]
hdr=None; cur={}
for r in rows:
    if r[0]=="ID": hdr=r; continue
    d=dict(zip(hdr,r))
    k=(d["ID"], d["Kernel Name"][:60])
    cur.setdefault(k,{})[d["Metric Name"]]=d["Metric Value"]
for k,v in cur.items(): print(k[0], k[1].split("(")[0][-40:], v)
PY
