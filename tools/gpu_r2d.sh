#!/bin/bash
# Round 2, GPU pass d: device-resident loop -- parity suite, config 1 / config 2 timings.
mkdir -p gpurun_out
tag=${1:-r2d}
timeout 300 python -m pytest tests/test_gpu_loop.py -x -q -p no:cacheprovider --timeout 120 > gpurun_out/${tag}_loop_tests.log 2>&1
echo "rc=$?" >> gpurun_out/${tag}_loop_tests.log
tail -30 gpurun_out/${tag}_loop_tests.log
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 -n 4 --durations=8 \
    > gpurun_out/${tag}_tests.log 2>&1
echo "rc=$?" >> gpurun_out/${tag}_tests.log
tail -25 gpurun_out/${tag}_tests.log
timeout 300 python tools/bench_small.py > gpurun_out/${tag}_config1.jsonl 2> gpurun_out/${tag}_config1.err
cut -c1-420 gpurun_out/${tag}_config1.jsonl; tail -3 gpurun_out/${tag}_config1.err
timeout 300 python tools/bench_configs.py 2 > gpurun_out/${tag}_config2.jsonl 2> gpurun_out/${tag}_config2.err
cut -c1-260 gpurun_out/${tag}_config2.jsonl
HP_B200_DEVICE_LOOP=0 timeout 300 python tools/bench_configs.py 2 > gpurun_out/${tag}_config2_hostloop.jsonl 2> gpurun_out/${tag}_config2_hostloop.err
cut -c1-260 gpurun_out/${tag}_config2_hostloop.jsonl
timeout 300 python tools/bench_configs.py 3 > gpurun_out/${tag}_config3.jsonl 2> gpurun_out/${tag}_config3.err
cut -c1-260 gpurun_out/${tag}_config3.jsonl
timeout 300 python bench.py --no-cpu-baseline --local-radius 0 --no-unscreened > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - "$tag" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/{sys.argv[1]}_bench.json"))
print("ms/step %.2f" % d["ms_per_step"], "kernel %.2f" % d["roofline"]["kernel_ms"], "frac %.4f" % d["roofline"]["frac"],
      "value %.3e e2e %.3e (%.3f s)" % (d["value"], d["e2e"]["value"], d["e2e"]["seconds"]), "hash", d["charges_sha256_10dec"])
PY
