#!/bin/bash
# A/B of a tuning build against the default library: parity suite on the default, then the N=1
# bench line (no CPU leg, no cut-off extra) for both.  usage: gpu_ab.sh <tag> <alt .so>
mkdir -p gpurun_out
tag=$1; alt=$2
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 120 -n 4 \
    > gpurun_out/${tag}_tests.log 2>&1
echo "rc=$?" >> gpurun_out/${tag}_tests.log
tail -8 gpurun_out/${tag}_tests.log
timeout 150 python bench.py --no-cpu-baseline --local-radius 0 > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err
HP_B200_LIB=$alt timeout 150 python bench.py --no-cpu-baseline --local-radius 0 > gpurun_out/${tag}_bench_alt.json 2> gpurun_out/${tag}_bench_alt.err
python - <<PY
import json
for n in ("default","alt"):
    try:
        d=json.load(open("gpurun_out/${tag}_bench_%s.json"%n))
        print(n, "ms/step %.2f"%d["ms_per_step"], "kernel %.2f"%d["roofline"]["kernel_ms"], "frac %.4f"%d["roofline"]["frac"],
              "unscreened %.2f ms frac %.4f"%(d["unscreened"]["kernel_ms"], d["unscreened"]["frac_algorithmic"]),
              "charges", d["charges_O_H_H"], "dq_unscr %.2e"%d["unscreened"]["max_abs_charge_diff_vs_screened"], "clk", d["clocks"]["sm_mhz"])
    except Exception as e:
        print(n, "failed", e)
PY
