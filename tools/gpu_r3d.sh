#!/bin/bash
mkdir -p gpurun_out
for v in sub1280 sub5120 sub1280p2; do
  echo "== $v"
  HP_B200_LIB=$PWD/horton_part_b200/libhp_${v}.so python tools/hessian_phases.py 2>&1 | grep "screened" | head -2
done | tee gpurun_out/r3d_hessian_variants.txt
