#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_glisa.py tests/test_gpu_schemes.py tests/test_gpu_configs.py tests/test_gpu_molgrid.py tests/test_gpu_solvers.py -q -x -m gpu 2>&1 | tail -4
python tools/bench_configs.py 4 > gpurun_out/r3s_config4.jsonl 2> gpurun_out/r3s.err
HP_B200_MOMENTS_SCREEN=0 python tools/bench_configs.py 4 > gpurun_out/r3s_config4_noscreen.jsonl 2>> gpurun_out/r3s.err
python - <<'EOF'
import json
for f in ("gpurun_out/r3s_config4.jsonl","gpurun_out/r3s_config4_noscreen.jsonl"):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d["seconds"], d["gradient_pass_ms"], d["roofline_hessian"]["ms"], d["charges_head"])
EOF
