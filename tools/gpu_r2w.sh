#!/bin/bash
mkdir -p gpurun_out
python tools/bench_configs.py 4 > gpurun_out/r2w_config4.jsonl 2> gpurun_out/r2w_config4.err
cut -c1-700 gpurun_out/r2w_config4.jsonl
python -m pytest tests -q -x -m gpu 2>&1 | tail -6 | tee gpurun_out/r2w_tests_full.txt
