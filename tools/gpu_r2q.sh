#!/bin/bash
mkdir -p gpurun_out
tag=${1:-r2q}
for v in b200 hs16x4 hs16x6 hs32x2; do
  HP_B200_LIB=$PWD/horton_part_b200/libhp_${v}.so timeout 600 python tools/bench_configs.py 4 > gpurun_out/${tag}_config4_${v}.jsonl 2> gpurun_out/${tag}_config4_${v}.err
  python - "$tag" "$v" <<'PY'
import json, sys
for line in open(f"gpurun_out/{sys.argv[1]}_config4_{sys.argv[2]}.jsonl"):
    d = json.loads(line)
    print(sys.argv[2], "config4 niter %d s/newton %.3f hessian %.1f ms %.2f TF frac %.3f" % (d["niter"], d["seconds_per_newton_iteration"], d["roofline_hessian"]["ms"], d["roofline_hessian"]["achieved"], d["roofline_hessian"]["frac"]))
PY
  tail -2 gpurun_out/${tag}_config4_${v}.err
done
