#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'syrk|basis_chunk|chunk_max|hessian_finish|tile_list' --csv --log-file gpurun_out/r3b_hessian_launches.csv python tools/hessian_once.py > gpurun_out/r3b_hessian_once.log 2>&1
python - <<'EOF'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r3b_hessian_launches.csv")) if len(r)>10 and r[0].isdigit()]
half=len(rows)//2
agg=collections.defaultdict(lambda:[0,0.0])
import re
for r in rows[half:]:
    name=r[4].split("(")[0]; v=float(r[-1].replace(",","")); unit=r[-2]
    if unit=="us": v/=1e3
    elif unit=="ns": v/=1e6
    elif unit=="s": v*=1e3
    agg[name][0]+=1; agg[name][1]+=v
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(f"{t:10.2f} ms x {n:5d}  {k}")
EOF
tail -2 gpurun_out/r3b_hessian_once.log
