#!/bin/bash
# ncu --set full of the screened Hessian kernels (config 4), a launch from the middle of the second Hessian
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"syrk_panel|basis_chunk" -s 120 -c 2 \
    -f -o gpurun_out/r3f_hessian_full python tools/hessian_once.py > gpurun_out/r3f_ncu_hessian.out 2>&1
tail -3 gpurun_out/r3f_ncu_hessian.out
ls -la gpurun_out/r3f_hessian_full.ncu-rep
