#!/bin/bash
# Parity suite, then the two ncu passes of /opt/skills/guides/B200_PROFILING.md on a 600-atom
# instance of the bench workload (numbers under a profiler are never bench values).
mkdir -p gpurun_out
tag=${1:-prof}
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 120 -n 4 > gpurun_out/${tag}_tests.log 2>&1
echo "rc=$?" >> gpurun_out/${tag}_tests.log
tail -6 gpurun_out/${tag}_tests.log
BENCH="python bench.py --natom 600 --steps 2 --warmup 1 --no-cpu-baseline --local-radius 0 --no-unscreened"
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv \
    --log-file gpurun_out/${tag}_launches.csv $BENCH > gpurun_out/${tag}_launches.out 2>&1
echo "launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:promol_weights_local -s 2 -c 1 \
    -f -o gpurun_out/${tag}_promol_full $BENCH > gpurun_out/${tag}_full.out 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | tail -8
