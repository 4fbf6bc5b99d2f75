#!/bin/bash
mkdir -p gpurun_out
for v in dk16s6 w44 w44dk16s6 w82; do
  echo "== $v"
  HP_B200_LIB=$PWD/horton_part_b200/libhp_${v}.so timeout 200 python tools/hessian_phases.py 2>&1 | grep "screened" | head -2
done | tee gpurun_out/r3o_hessian_variants.txt
