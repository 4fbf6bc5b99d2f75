#!/bin/bash
# Round 2, GPU pass c: parity suite on the new defaults (one-constant exp, 3 blocks/SM, LUT spline kernel),
# config 2 timings, bench, ncu captures of the hot kernel / spline kernel / aLISA block solver.
mkdir -p gpurun_out
tag=${1:-r2c}
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 -n 4 --durations=10 \
    > gpurun_out/${tag}_tests.log 2>&1
echo "rc=$?" >> gpurun_out/${tag}_tests.log
tail -25 gpurun_out/${tag}_tests.log
timeout 300 python tools/bench_configs.py 2 > gpurun_out/${tag}_config2.jsonl 2> gpurun_out/${tag}_config2.err
cut -c1-260 gpurun_out/${tag}_config2.jsonl
timeout 300 python bench.py --no-cpu-baseline --local-radius 0 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - "$tag" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/{sys.argv[1]}_bench.json"))
print("ms/step %.2f" % d["ms_per_step"], "kernel %.2f" % d["roofline"]["kernel_ms"], "frac %.4f" % d["roofline"]["frac"],
      "unscreened %.2f ms frac %.4f" % (d["unscreened"]["kernel_ms"], d["unscreened"]["frac_executed"]),
      "dq_unscr %.2e" % d["unscreened"]["max_abs_charge_diff_vs_screened"], "value %.3e e2e %.3e" % (d["value"], d["e2e"]["value"]),
      "hash", d["charges_sha256_10dec"], d["charges_O_H_H"])
PY
BENCH="python bench.py --natom 600 --steps 2 --warmup 1 --no-cpu-baseline --local-radius 0 --no-unscreened"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:promol_weights_local -s 2 -c 1 \
    -f -o gpurun_out/${tag}_promol_full $BENCH > gpurun_out/${tag}_ncu_promol.out 2>&1
echo "hot kernel capture rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:promol_weights_spline -s 3 -c 1 \
    -f -o gpurun_out/${tag}_spline_full python tools/bench_configs.py 2 > gpurun_out/${tag}_ncu_spline.out 2>&1
echo "spline capture rc=$?"
ls -la gpurun_out | grep ${tag}
