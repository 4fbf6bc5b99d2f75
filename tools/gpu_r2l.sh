#!/bin/bash
mkdir -p gpurun_out
tag=${1:-r2l}
timeout 300 python -m pytest tests/test_gpu_schemes.py tests/test_gpu_glisa.py tests/test_gpu_configs.py -q -p no:cacheprovider --timeout 300 -x > gpurun_out/${tag}_new_tests.log 2>&1
echo "rc=$?" >> gpurun_out/${tag}_new_tests.log
tail -30 gpurun_out/${tag}_new_tests.log
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 -n 4 > gpurun_out/${tag}_tests.log 2>&1
echo "rc=$?" >> gpurun_out/${tag}_tests.log
tail -8 gpurun_out/${tag}_tests.log
timeout 600 python tools/bench_configs.py 4 > gpurun_out/${tag}_configs.jsonl 2> gpurun_out/${tag}_configs.err
python - "$tag" <<'PY'
import json, sys
for line in open(f"gpurun_out/{sys.argv[1]}_configs.jsonl"):
    d = json.loads(line)
    print("config4 niter %d s/newton %.3f hessian %.1f ms %.2f TF frac %.3f grad %.1f ms" % (d["niter"], d["seconds_per_newton_iteration"], d["roofline_hessian"]["ms"], d["roofline_hessian"]["achieved"], d["roofline_hessian"]["frac"], d["gradient_pass_ms"]), d["charges_head"])
PY
tail -3 gpurun_out/${tag}_configs.err
