#!/bin/bash
mkdir -p gpurun_out
for v in b200 rows16 rows64 rows32b4; do
  echo "== $v"
  HP_B200_LIB=$PWD/horton_part_b200/libhp_${v}.so python tools/hessian_phases.py 2>&1 | grep "screened" | head -2
done | tee gpurun_out/r3g_hessian_variants.txt
bash tools/gpu_r3b.sh
