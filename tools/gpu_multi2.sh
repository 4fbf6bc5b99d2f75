#!/bin/bash
# Two-GPU check: the sharded parity tests and the N=2 bench line (torchrun, NCCL over NVLink).
mkdir -p gpurun_out
tag=${1:-multi2}
timeout 240 python -m pytest tests/test_gpu_multi.py -q -p no:cacheprovider --timeout 200 > gpurun_out/${tag}_tests.log 2>&1
echo "rc=$?" >> gpurun_out/${tag}_tests.log
tail -6 gpurun_out/${tag}_tests.log
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --no-cpu-baseline > gpurun_out/${tag}_bench_n2.json 2> gpurun_out/${tag}_bench_n2.err
echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench_n2.json"))
print("n_gpus",d["n_gpus"],"ms/step %.2f"%d["ms_per_step"],"value %.4g"%d["value"],"e2e %.4g"%d["e2e"]["value"],"unscreened %.2f"%d["unscreened"]["ms_per_step"])
PY
