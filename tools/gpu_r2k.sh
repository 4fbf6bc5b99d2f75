#!/bin/bash
mkdir -p gpurun_out
tag=${1:-r2k}
timeout 600 python tools/e2e_profile_sharded.py > gpurun_out/${tag}_e2e_sharded.txt 2>&1; head -70 gpurun_out/${tag}_e2e_sharded.txt | cut -c1-180
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"syrk_panel_dmma|basis_chunk" -s 4 -c 2 \
    -f -o gpurun_out/${tag}_hessian_full python tools/bench_configs.py 4 > gpurun_out/${tag}_ncu_hessian.out 2>&1
echo "hessian capture rc=$?"
