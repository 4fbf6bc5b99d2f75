#!/bin/bash
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r4b_scale_n1.json 2> gpurun_out/r4b_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 8 --steps 20 --warmup 3 --no-extras > gpurun_out/r4b_scale_n8.json 2> gpurun_out/r4b_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 8 --steps 5 --warmup 3 --no-extras > gpurun_out/r4b_bench_n8_steps5.json 2>> gpurun_out/r4b_n8.err
python - <<'EOF'
import json
for f in ("r4b_scale_n1","r4b_scale_n8","r4b_bench_n8_steps5"):
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    e=d["e2e"]
    print(f, d["n_gpus"], d["steps"], "ms/step %.2f"%d["ms_per_step"], "value %.3e"%d["value"], "frac %.3f"%d["roofline"]["frac"], "e2e %.3f s (pageable %.3f)"%(e["seconds"], e["pageable_inputs"]["seconds"]), "e2e value %.3e"%e["value"], d["charges_sha256_10dec"])
EOF
