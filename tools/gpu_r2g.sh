#!/bin/bash
mkdir -p gpurun_out
tag=${1:-r2g}
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 -n 4 > gpurun_out/${tag}_tests.log 2>&1
echo "rc=$?" >> gpurun_out/${tag}_tests.log
tail -30 gpurun_out/${tag}_tests.log
timeout 600 python tools/bench_configs.py 2 4 > gpurun_out/${tag}_configs.jsonl 2> gpurun_out/${tag}_configs.err
python - "$tag" <<'PY'
import json, sys
for line in open(f"gpurun_out/{sys.argv[1]}_configs.jsonl"):
    d = json.loads(line)
    if d["config"] == 2:
        print("config2 isa", d["isa"]["ms_per_iteration"], d["isa"]["gpu_ms_weights_per_iteration"], d["isa"]["roofline"]["kernel_ms"], d["isa"]["roofline"]["frac"], "mbis", d["mbis"]["ms_per_iteration"])
    else:
        print("config4", d["niter"], d["seconds_per_newton_iteration"], d["roofline_hessian"]["ms"], d["roofline_hessian"]["achieved"], d["roofline_hessian"]["frac"])
PY
tail -3 gpurun_out/${tag}_configs.err
