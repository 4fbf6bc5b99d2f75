#!/bin/bash
mkdir -p gpurun_out
tag=${1:-r2i}
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 -n 4 > gpurun_out/${tag}_tests.log 2>&1
echo "rc=$?" >> gpurun_out/${tag}_tests.log
tail -12 gpurun_out/${tag}_tests.log
timeout 600 python tools/bench_configs.py 1 2 3 > gpurun_out/${tag}_configs.jsonl 2> gpurun_out/${tag}_configs.err
python - "$tag" <<'PY'
import json, sys
for line in open(f"gpurun_out/{sys.argv[1]}_configs.jsonl"):
    d = json.loads(line)
    c = d["config"]
    if c == 1:
        print("config1 e2e best %.2f ms gpu %.2f ms" % (1e3 * d["e2e_seconds_best"], 1e3 * d["gpu_seconds"]))
    elif c == 2:
        print("config2 isa ms/it %.3f weights %.3f propars %.3f kernel %.3f frac %.3f | mbis ms/it %.3f w %.3f p %.3f" % (d["isa"]["ms_per_iteration"], d["isa"]["gpu_ms_weights_per_iteration"], d["isa"]["gpu_ms_propars_per_iteration"], d["isa"]["roofline"]["kernel_ms"], d["isa"]["roofline"]["frac"], d["mbis"]["ms_per_iteration"], d["mbis"]["gpu_ms_weights_per_iteration"], d["mbis"]["gpu_ms_propars_per_iteration"]))
    elif c == 3:
        for b in ("gauss", "slater"):
            r = d[b]
            print("config3", b, "ms/it %.3f weights %.3f (frac %.3f) radial %.3f" % (r["ms_per_iteration"], r["roofline_weights"]["kernel_ms"], r["roofline_weights"]["frac"], r["radial_solver"]["ms_per_iteration"]))
PY
tail -3 gpurun_out/${tag}_configs.err
timeout 300 python tools/e2e_profile.py > gpurun_out/${tag}_e2e_profile.txt 2>&1; tail -16 gpurun_out/${tag}_e2e_profile.txt
