#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"molgrid_reduce" -s 2 -c 1 \
    -f -o gpurun_out/r4k_moments_full python tools/bench_configs.py 4 > gpurun_out/r4k_ncu.out 2>&1
ls -la gpurun_out/r4k_moments_full.ncu-rep
python bench.py > gpurun_out/r4k_bench.json 2> gpurun_out/r4k_bench.err
tail -2 gpurun_out/r4k_bench.err
python - <<'EOF'
import json
d=json.loads(open("gpurun_out/r4k_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches","charges_sha256_10dec")}, d["roofline"]["frac"], d["e2e"]["seconds"], d["e2e"]["pageable_inputs"]["seconds"])
c=d["configs_1_to_4"]
print("c1", c["config1"]["e2e_seconds_best"], "c2 isa", c["config2"]["isa"]["ms_per_iteration"], "hi", c["config2"]["hirshfeld_i"]["seconds"], "c3", c["config3"]["gauss"]["ms_per_iteration"], c["config3"]["slater"]["ms_per_iteration"])
c4=c["config4"]; print("c4", c4["seconds"], c4["roofline_hessian"]["ms"], c4["roofline_hessian"]["frac"], c4["roofline_hessian_unscreened"]["ms"], c4["roofline_hessian_unscreened"]["achieved"], c4["gradient_pass_ms"])
EOF
