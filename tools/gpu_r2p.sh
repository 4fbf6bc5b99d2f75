#!/bin/bash
mkdir -p gpurun_out
tag=${1:-r2p}
timeout 300 python -m pytest tests/test_gpu_loop.py tests/test_gpu_isa.py tests/test_gpu_glisa.py -q -p no:cacheprovider --timeout 300 > gpurun_out/${tag}_tests.log 2>&1
echo "rc=$?" >> gpurun_out/${tag}_tests.log; tail -4 gpurun_out/${tag}_tests.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"becke_weights|lisa_sc_block|molgrid_update" -c 5 \
    -f -o gpurun_out/${tag}_misc_full python tools/ncu_misc.py > gpurun_out/${tag}_ncu_misc.out 2>&1
echo "misc capture rc=$?"; tail -4 gpurun_out/${tag}_ncu_misc.out
