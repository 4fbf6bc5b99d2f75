#!/bin/bash
# Hessian block screening: parity tests + config-4 timing, screened and unscreened
mkdir -p gpurun_out
python -m pytest tests/test_gpu_glisa.py tests/test_gpu_configs.py tests/test_gpu_schemes.py -q -x -m gpu 2>&1 | tail -5 > gpurun_out/r2r_tests.txt
python tools/bench_configs.py 4 > gpurun_out/r2r_config4.jsonl 2> gpurun_out/r2r_config4.err
HP_B200_HESSIAN_SCREEN=0 python tools/bench_configs.py 4 > gpurun_out/r2r_config4_noscreen.jsonl 2>> gpurun_out/r2r_config4.err
cat gpurun_out/r2r_tests.txt
tail -c 1500 gpurun_out/r2r_config4.jsonl
tail -c 600 gpurun_out/r2r_config4_noscreen.jsonl
