#!/bin/bash
# One bounded GPU validation pass (run through gpurun): new/changed tests first, then the whole
# -m gpu suite on 4 xdist workers, then the N=1 bench line.  Logs land in gpurun_out/.
mkdir -p gpurun_out
tag=${1:-check}
python -c "import torch; print(torch.cuda.get_device_name(0))" > gpurun_out/${tag}_env.log 2>&1
timeout 200 python -m pytest -q -p no:cacheprovider --timeout 120 \
    tests/test_gpu_solvers.py -k "convex" \
    tests/test_gpu_molgrid.py \
    tests/test_gpu_schemes.py tests/test_gpu_hirshfeld.py tests/test_gpu_ragged.py \
    > gpurun_out/${tag}_new_tests.log 2>&1
echo "rc=$?" >> gpurun_out/${tag}_new_tests.log
tail -25 gpurun_out/${tag}_new_tests.log
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 120 -n 4 --durations=15 \
    > gpurun_out/${tag}_all_tests.log 2>&1
echo "rc=$?" >> gpurun_out/${tag}_all_tests.log
tail -30 gpurun_out/${tag}_all_tests.log
timeout 200 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
echo "bench rc=$?"
head -c 600 gpurun_out/${tag}_bench_n1.json
