#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_hostmem.py tests/test_gpu_mbis.py tests/test_gpu_glisa.py tests/test_gpu_local.py tests/test_gpu_screening.py tests/test_gpu_loop.py -q -x -m gpu 2>&1 | tail -5
timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r3u_bench.json 2> gpurun_out/r3u_bench.err
tail -2 gpurun_out/r3u_bench.err
python - <<'EOF'
import json
d=json.loads(open("gpurun_out/r3u_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","charges_sha256_10dec")}, d["roofline"]["frac"])
e=d["e2e"]; print("e2e pinned", e["seconds"], "first", e["seconds_first_call"], "pageable", e["pageable_inputs"]["seconds"], e["hostmem"])
EOF
HP_B200_SPLIT_UPLOAD=0 timeout 300 python bench.py --no-extras --no-cpu-baseline --no-unscreened --local-radius 0 > gpurun_out/r3u_bench_nosplit.json 2>> gpurun_out/r3u_bench.err
python - <<'EOF'
import json
d=json.loads(open("gpurun_out/r3u_bench_nosplit.json").read().strip().splitlines()[-1])
e=d["e2e"]; print("nosplit: e2e pinned", e["seconds"], "pageable", e["pageable_inputs"]["seconds"], d["charges_sha256_10dec"])
EOF
