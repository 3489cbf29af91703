#!/bin/bash
# round-2 call 5: where do the chain kernels spend their time?  (ncu --set full on every nbk_ kernel of one step)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nb_ops.py tests/test_gpu_chain.py -m gpu -q > gpurun_out/c5_tests_nb.log 2>&1
echo "nb tests exit $?" >> gpurun_out/c5_tests_nb.log
timeout 300 python tools/nb_bench.py > gpurun_out/c5_nb_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nbk_ -c 40 -o gpurun_out/c5_nbk_prof python bench.py --steps 1 --warmup 1 --no-gpu-reference --no-roofline --no-cpu-baseline --no-graph > gpurun_out/c5_ncu_nbk.log 2>&1
tail -12 gpurun_out/c5_tests_nb.log; cat gpurun_out/c5_nb_bench.log; tail -3 gpurun_out/c5_ncu_nbk.log
