#!/bin/bash
# Round 2, GPU pass f: full parity suite on everything new (block radial solvers, DMMA Hessian, parallel spline
# build, device cube arrays, grid_type 3 integrals), measured-deviation report, configs 1-4 blocks, bench line.
mkdir -p gpurun_out
tag=${1:-r2f}
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 -n 4 --durations=8 \
    > gpurun_out/${tag}_tests.log 2>&1
echo "rc=$?" >> gpurun_out/${tag}_tests.log
tail -40 gpurun_out/${tag}_tests.log
timeout 600 python tools/parity_report.py > gpurun_out/${tag}_parity_report.txt 2> gpurun_out/${tag}_parity_report.err
cat gpurun_out/${tag}_parity_report.txt; tail -5 gpurun_out/${tag}_parity_report.err
timeout 600 python tools/bench_configs.py 1 2 3 4 > gpurun_out/${tag}_configs.jsonl 2> gpurun_out/${tag}_configs.err
cut -c1-1500 gpurun_out/${tag}_configs.jsonl; tail -5 gpurun_out/${tag}_configs.err
HP_B200_HESSIAN_DFMA=1 timeout 300 python tools/bench_configs.py 4 > gpurun_out/${tag}_config4_dfma.jsonl 2> gpurun_out/${tag}_config4_dfma.err
cut -c1-900 gpurun_out/${tag}_config4_dfma.jsonl
timeout 600 python bench.py --no-extras > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - "$tag" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/{sys.argv[1]}_bench.json"))
print("ms/step %.2f" % d["ms_per_step"], "kernel %.2f" % d["roofline"]["kernel_ms"], "frac %.4f" % d["roofline"]["frac"],
      "value %.3e job %.3e e2e %.3e (%.3f s)" % (d["value"], d["value_job"], d["e2e"]["value"], d["e2e"]["seconds"]),
      "cpu %.3e x%d" % (d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"]), "hash", d["charges_sha256_10dec"])
PY
