#!/bin/bash
# Round 2, GPU pass e: FP64 tensor-core probe + cuBLAS yardstick for the Hessian decision, per-kernel launch
# lists of the small configurations (host-driven loop so that every launch is visible to ncu).
mkdir -p gpurun_out
tag=${1:-r2e}
nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/dmma_probe.cu -o /tmp/dmma && timeout 120 /tmp/dmma > gpurun_out/${tag}_dmma_probe.txt 2>&1
cat gpurun_out/${tag}_dmma_probe.txt
timeout 120 python tools/dgemm_yardstick.py > gpurun_out/${tag}_dgemm.json 2>&1; cat gpurun_out/${tag}_dgemm.json
timeout 300 python -m pytest tests/test_gpu_loop.py -q -p no:cacheprovider --timeout 120 > gpurun_out/${tag}_loop_tests.log 2>&1
tail -3 gpurun_out/${tag}_loop_tests.log
export HP_B200_DEVICE_LOOP=0
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 80 --csv \
    --log-file gpurun_out/${tag}_launches_config1.csv python tools/bench_small.py > gpurun_out/${tag}_l1.out 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv \
    --log-file gpurun_out/${tag}_launches_config2.csv python tools/bench_configs.py 2 > gpurun_out/${tag}_l2.out 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 60 --csv \
    --log-file gpurun_out/${tag}_launches_config3.csv python tools/bench_configs.py 3 > gpurun_out/${tag}_l3.out 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 300 --csv \
    --log-file gpurun_out/${tag}_launches_config4.csv python tools/bench_configs.py 4 > gpurun_out/${tag}_l4.out 2>&1
python - "$tag" <<'PY'
import csv, sys, collections
tag = sys.argv[1]
for cfg in ("config1", "config2", "config3", "config4"):
    try:
        rows = [r for r in csv.reader(open(f"gpurun_out/{tag}_launches_{cfg}.csv")) if len(r) > 5 and r[0].isdigit()]
    except Exception as e:
        print(cfg, "failed", e); continue
    agg = collections.OrderedDict()
    for r in rows:
        name = r[4].split("(")[0][-60:]
        val = float(r[-1].replace(",", "")) ; unit = r[-2]
        us = val / 1e3 if unit in ("ns", "nsecond") else (val if unit.startswith("us") else val * 1e3)
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us
    tot = sum(v[1] for v in agg.values())
    print(f"== {cfg}: {len(rows)} launches, {tot:.0f} us")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
        print(f"   {us/tot*100:5.1f}%  {us/n:9.1f} us x {n:4d}  {k}")
PY
