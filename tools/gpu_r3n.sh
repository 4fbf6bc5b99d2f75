#!/bin/bash
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_glisa.py -q -x -m gpu 2>&1 | tail -3
echo "rc=$?"
echo "== bulk"; timeout 200 python tools/hessian_phases.py 2>&1 | grep "screened" | head -2
echo "== cpasync"; HP_B200_HESSIAN_PIPE=cpasync timeout 200 python tools/hessian_phases.py 2>&1 | grep "screened" | head -2
