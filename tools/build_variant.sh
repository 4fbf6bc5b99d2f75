#!/bin/bash
# Build a tuning variant of the library next to the default one: tools/build_variant.sh <name> <nvcc -D flags...>
# -> horton_part_b200/libhp_<name>.so (git-ignored; select it with HP_B200_LIB=<path>)
set -e
name=$1; shift
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
    --shared -cudart shared -I include -I horton_part_b200/csrc "$@" horton_part_b200/csrc/*.cu -o horton_part_b200/libhp_${name}.so -ldl
echo horton_part_b200/libhp_${name}.so
