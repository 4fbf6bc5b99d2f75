#!/bin/bash
mkdir -p gpurun_out
python tools/hessian_phases.py 2>&1 | grep "screened" | tee gpurun_out/r3c_hessian.txt
bash tools/gpu_r3b.sh
python -m pytest tests/test_gpu_glisa.py tests/test_gpu_configs.py tests/test_gpu_schemes.py -q -x -m gpu 2>&1 | tail -3
