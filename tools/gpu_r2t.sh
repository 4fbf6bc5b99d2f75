#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_glisa.py tests/test_gpu_configs.py tests/test_gpu_schemes.py -q -x -m gpu 2>&1 | tail -5 > gpurun_out/r2t_tests.txt
python tools/bench_configs.py 4 > gpurun_out/r2t_config4.jsonl 2> gpurun_out/r2t_config4.err
HP_B200_HOST_SOLVE=1 python tools/bench_configs.py 4 > gpurun_out/r2t_config4_hostsolve.jsonl 2>> gpurun_out/r2t_config4.err
cat gpurun_out/r2t_tests.txt
cut -c1-420 gpurun_out/r2t_config4.jsonl
cut -c1-420 gpurun_out/r2t_config4_hostsolve.jsonl
