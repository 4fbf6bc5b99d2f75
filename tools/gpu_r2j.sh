#!/bin/bash
# 2-GPU pass: sharded parity tests (torch.distributed NCCL group and the C-ABI communicator), bench at N=2.
mkdir -p gpurun_out
tag=${1:-r2j}
nvidia-smi -L > gpurun_out/${tag}_env.log 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -q -p no:cacheprovider --timeout 300 > gpurun_out/${tag}_multi_tests.log 2>&1
echo "rc=$?" >> gpurun_out/${tag}_multi_tests.log
tail -15 gpurun_out/${tag}_multi_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --no-extras > gpurun_out/${tag}_bench_n2.json 2> gpurun_out/${tag}_bench_n2.err
python - "$tag" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/{sys.argv[1]}_bench_n2.json"))
print("N=2 ms/step %.2f" % d["ms_per_step"], "kernel %.2f" % d["roofline"]["kernel_ms"], "frac %.4f" % d["roofline"]["frac"],
      "value %.3e job %.3e e2e %.3e (%.3f s)" % (d["value"], d["value_job"], d["e2e"]["value"], d["e2e"]["seconds"]), "hash", d["charges_sha256_10dec"])
PY
tail -3 gpurun_out/${tag}_bench_n2.err
# config 3 (aLISA, 100 atoms) on 2 GPUs, as BASELINE.json asks
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
    tools/bench_config3_multi.py > gpurun_out/${tag}_config3_n2.jsonl 2> gpurun_out/${tag}_config3_n2.err
cat gpurun_out/${tag}_config3_n2.jsonl; tail -3 gpurun_out/${tag}_config3_n2.err
