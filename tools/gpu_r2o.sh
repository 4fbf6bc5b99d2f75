#!/bin/bash
# 8-GPU pass: bench at N = 8 and 4 (strong scaling of config 5), configs 3 and 4 sharded over 8 GPUs.
mkdir -p gpurun_out
tag=${1:-r2o}
nvidia-smi -L > gpurun_out/${tag}_env.log 2>&1
for n in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
      bench.py --gpus $n --steps 20 --warmup 3 --no-extras > gpurun_out/${tag}_bench_n$n.json 2> gpurun_out/${tag}_bench_n$n.err
  python - "$tag" "$n" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/{sys.argv[1]}_bench_n{sys.argv[2]}.json"))
print("N=%s ms/step %.2f" % (sys.argv[2], d["ms_per_step"]), "kernel %.2f" % d["roofline"]["kernel_ms"], "frac %.4f" % d["roofline"]["frac"],
      "value %.3e job %.3e e2e %.3e (%.3f s)" % (d["value"], d["value_job"], d["e2e"]["value"], d["e2e"]["seconds"]), "hash", d["charges_sha256_10dec"], d["clocks"])
PY
  tail -2 gpurun_out/${tag}_bench_n$n.err
done
timeout 600 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
python - "$tag" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/{sys.argv[1]}_bench_n1.json"))
print("N=1 ms/step %.2f" % d["ms_per_step"], "value %.3e job %.3e e2e %.3e (%.3f s)" % (d["value"], d["value_job"], d["e2e"]["value"], d["e2e"]["seconds"]), "hash", d["charges_sha256_10dec"])
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
    tools/bench_config3_multi.py 3 4 > gpurun_out/${tag}_configs34_n8.jsonl 2> gpurun_out/${tag}_configs34_n8.err
cat gpurun_out/${tag}_configs34_n8.jsonl; tail -3 gpurun_out/${tag}_configs34_n8.err
