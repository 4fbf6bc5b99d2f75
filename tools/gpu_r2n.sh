#!/bin/bash
mkdir -p gpurun_out
tag=${1:-r2n}
( time timeout 900 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err ) 2> gpurun_out/${tag}_bench_time.txt
echo "bench rc=$?"; cat gpurun_out/${tag}_bench_time.txt | tail -4
python - "$tag" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/{sys.argv[1]}_bench_n1.json"))
print("ms/step %.2f" % d["ms_per_step"], "kernel %.2f" % d["roofline"]["kernel_ms"], "frac %.4f" % d["roofline"]["frac"],
      "value %.3e job %.3e e2e %.3e (%.3f s)" % (d["value"], d["value_job"], d["e2e"]["value"], d["e2e"]["seconds"]),
      "traffic", d["roofline"]["traffic"], d["roofline"].get("traffic_over_algorithmic"), "hash", d["charges_sha256_10dec"], "clocks", d["clocks"])
ex = d.get("configs_1_to_4", {})
for k, v in ex.items():
    print(k, "ERROR " + v["error"] if "error" in v else "ok", str(v)[:300])
print("cpu", d.get("cpu_baseline"))
PY
( time timeout 600 python bench.py --impl reference > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err ) 2>&1 | tail -4
cut -c1-900 gpurun_out/${tag}_bench_ref.json
python -c "
import __graft_entry__ as g
g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -3 gpurun_out/${tag}_smoke.log
