#!/bin/bash
# Round 2, first GPU pass: full parity suite, FP64 instruction-mix probe, config 3 with the block-per-atom
# radial solver vs the one-warp kernel, config-5 bench default vs base-2 exponential build.
mkdir -p gpurun_out
tag=${1:-r2a}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_env.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 -n 4 --durations=20 \
    > gpurun_out/${tag}_tests.log 2>&1
echo "rc=$?" >> gpurun_out/${tag}_tests.log
tail -40 gpurun_out/${tag}_tests.log
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I horton_part_b200/csrc -I include tools/fp64_mix_probe.cu -o /tmp/mix \
    && timeout 120 /tmp/mix > gpurun_out/${tag}_mix_probe.txt 2>&1
cat gpurun_out/${tag}_mix_probe.txt
timeout 300 python tools/bench_configs.py 3 > gpurun_out/${tag}_config3_block.jsonl 2> gpurun_out/${tag}_config3_block.err
HP_B200_SC_GENERIC=1 timeout 300 python tools/bench_configs.py 3 > gpurun_out/${tag}_config3_generic.jsonl 2> gpurun_out/${tag}_config3_generic.err
cut -c1-300 gpurun_out/${tag}_config3_block.jsonl gpurun_out/${tag}_config3_generic.jsonl
timeout 300 python bench.py --no-cpu-baseline --local-radius 0 > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err
HP_B200_LIB=$PWD/horton_part_b200/libhp_exp2.so timeout 300 python bench.py --no-cpu-baseline --local-radius 0 \
    > gpurun_out/${tag}_bench_exp2.json 2> gpurun_out/${tag}_bench_exp2.err
python - <<PY
import json
for n in ("default","exp2"):
    try:
        d=json.load(open("gpurun_out/${tag}_bench_%s.json"%n))
        print(n, "ms/step %.2f"%d["ms_per_step"], "kernel %.2f"%d["roofline"]["kernel_ms"], "frac %.4f"%d["roofline"]["frac"],
              "unscreened %.2f ms frac %.4f"%(d["unscreened"]["kernel_ms"], d["unscreened"]["frac_executed"]),
              "charges", d["charges_O_H_H"], "dq_unscr %.2e"%d["unscreened"]["max_abs_charge_diff_vs_screened"], "clk", d["clocks"]["sm_mhz"],
              "value %.3e job %.3e e2e %.3e"%(d["value"], d["value_job"], d["e2e"]["value"]))
    except Exception as e:
        print(n, "failed", e)
PY
