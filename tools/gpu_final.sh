#!/bin/bash
# Round-end style pass on one GPU: the parity suite in ONE process (as the driver runs it), smoke(),
# the N=1 bench line, config 3 timings.
mkdir -p gpurun_out
tag=${1:-final}
timeout 400 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 150 > gpurun_out/${tag}_tests.log 2>&1
echo "rc=$?" >> gpurun_out/${tag}_tests.log
tail -5 gpurun_out/${tag}_tests.log
timeout 120 python __graft_entry__.py --smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${tag}_smoke.log
timeout 240 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; echo "bench rc=$?"
timeout 200 python tools/bench_configs.py 3 > gpurun_out/${tag}_config3.jsonl 2> gpurun_out/${tag}_config3.err; echo "config3 rc=$?"
cat gpurun_out/${tag}_config3.jsonl | cut -c1-260
