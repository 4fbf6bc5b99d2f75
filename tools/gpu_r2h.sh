#!/bin/bash
mkdir -p gpurun_out
tag=${1:-r2h}
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 -n 4 > gpurun_out/${tag}_tests.log 2>&1
echo "rc=$?" >> gpurun_out/${tag}_tests.log
tail -12 gpurun_out/${tag}_tests.log
timeout 600 python tools/bench_configs.py 1 2 3 4 > gpurun_out/${tag}_configs.jsonl 2> gpurun_out/${tag}_configs.err
python - "$tag" <<'PY'
import json, sys
for line in open(f"gpurun_out/{sys.argv[1]}_configs.jsonl"):
    d = json.loads(line)
    c = d["config"]
    if c == 1:
        print("config1 e2e best %.2f ms gpu %.2f ms" % (1e3 * d["e2e_seconds_best"], 1e3 * d["gpu_seconds"]))
    elif c == 2:
        print("config2 isa ms/it %.3f weights %.3f kernel %.3f frac %.3f | mbis ms/it %.3f" % (d["isa"]["ms_per_iteration"], d["isa"]["gpu_ms_weights_per_iteration"], d["isa"]["roofline"]["kernel_ms"], d["isa"]["roofline"]["frac"], d["mbis"]["ms_per_iteration"]))
    elif c == 3:
        for b in ("gauss", "slater"):
            r = d[b]
            print("config3", b, "ms/it %.3f weights %.3f (frac %.3f) radial %.3f" % (r["ms_per_iteration"], r["roofline_weights"]["kernel_ms"], r["roofline_weights"]["frac"], r["radial_solver"]["ms_per_iteration"]))
    else:
        print("config4 niter %d s/newton %.3f hessian %.1f ms %.2f TF frac %.3f" % (d["niter"], d["seconds_per_newton_iteration"], d["roofline_hessian"]["ms"], d["roofline_hessian"]["achieved"], d["roofline_hessian"]["frac"]))
PY
tail -3 gpurun_out/${tag}_configs.err
# ncu: launch list + full capture of the hot kernel on the bench workload itself (2,000 atoms)
BENCH="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --local-radius 0 --no-unscreened --no-extras"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/${tag}_launches_config5.csv $BENCH > gpurun_out/${tag}_l5.out 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:promol_weights_local -s 2 -c 1 \
    -f -o gpurun_out/${tag}_promol_full_config5 $BENCH > gpurun_out/${tag}_ncu5.out 2>&1
echo "full capture rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"promol_weights_spline|spline_build_inv|lisa_sc_block|shell_fixed_point|syrk_panel_dmma|basis_chunk" -c 12 \
    -f -o gpurun_out/${tag}_others_full python tools/bench_configs.py 2 > gpurun_out/${tag}_ncu_others.out 2>&1
echo "other kernels capture rc=$?"
ls -la gpurun_out | grep ${tag} | awk '{print $5, $9}'
