#!/bin/bash
# A/B of tuning builds of the library (tools/build_variant.sh): the config-5 bench line for each,
# optionally the parity suite on the first one.  usage: gpu_variants.sh <tag> <test-variant|-> <variant>...
mkdir -p gpurun_out
tag=$1; testv=$2; shift 2
if [ -x /tmp/mix ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I horton_part_b200/csrc -I include tools/fp64_mix_probe.cu -o /tmp/mix; then
    timeout 120 /tmp/mix > gpurun_out/${tag}_mix_probe.txt 2>&1; tail -12 gpurun_out/${tag}_mix_probe.txt
fi
if [ "$testv" != "-" ]; then
    HP_B200_LIB=$PWD/horton_part_b200/libhp_${testv}.so timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider \
        --timeout 600 -n 4 > gpurun_out/${tag}_tests_${testv}.log 2>&1
    echo "rc=$?" >> gpurun_out/${tag}_tests_${testv}.log
    tail -5 gpurun_out/${tag}_tests_${testv}.log
fi
for v in "$@"; do
    HP_B200_LIB=$PWD/horton_part_b200/libhp_${v}.so timeout 300 python bench.py --no-cpu-baseline --local-radius 0 \
        > gpurun_out/${tag}_bench_${v}.json 2> gpurun_out/${tag}_bench_${v}.err
done
python - "$tag" "$@" <<'PY'
import json, sys
tag = sys.argv[1]
for n in sys.argv[2:]:
    try:
        d = json.load(open(f"gpurun_out/{tag}_bench_{n}.json"))
        print(n, "ms/step %.2f" % d["ms_per_step"], "kernel %.2f" % d["roofline"]["kernel_ms"], "frac %.4f" % d["roofline"]["frac"],
              "unscreened %.2f ms frac %.4f" % (d["unscreened"]["kernel_ms"], d["unscreened"]["frac_executed"]),
              "pairs %.4f" % d["roofline"]["pairs_evaluated_fraction"], "dq_unscr %.2e" % d["unscreened"]["max_abs_charge_diff_vs_screened"],
              "clk", d["clocks"]["sm_mhz"], "value %.3e e2e %.3e" % (d["value"], d["e2e"]["value"]), "hash", d["charges_sha256_10dec"])
    except Exception as e:
        print(n, "failed", e)
PY
