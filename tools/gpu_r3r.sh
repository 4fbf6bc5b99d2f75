#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_r3b.sh > gpurun_out/r3r_hessian_launch_summary.txt 2>&1
cat gpurun_out/r3r_hessian_launch_summary.txt
python bench.py > gpurun_out/r3r_bench.json 2> gpurun_out/r3r_bench.err
tail -2 gpurun_out/r3r_bench.err
python - <<'EOF'
import json
d=json.loads(open("gpurun_out/r3r_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches","charges_sha256_10dec")}, d["roofline"]["frac"], d["e2e"]["seconds"], d["e2e"]["pageable_inputs"]["seconds"])
print(json.dumps(d["configs_1_to_4"]["config4"])[:900])
EOF
