#!/bin/bash
# round-2 call 3: rewritten tail kernels (no shared atomics, bulk-copy ring), optimizer-driven packs
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_chain.py -m gpu -q > gpurun_out/c3_tests_chain.log 2>&1
echo "chain tests exit $?" >> gpurun_out/c3_tests_chain.log
timeout 900 python -m pytest tests/test_gpu_tail.py tests/test_gpu_z_optimizer.py tests/test_gpu_dcgan.py tests/test_gpu_ops.py tests/test_gpu_models.py -m gpu -q > gpurun_out/c3_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/c3_tests.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tail_ -c 6 -o gpurun_out/c3_tail_prof python bench.py --steps 2 --warmup 1 --no-gpu-reference --no-cpu-baseline > gpurun_out/c3_ncu_tail.log 2>&1
tail -15 gpurun_out/c3_tests_chain.log; tail -3 gpurun_out/c3_tests.log; cut -c1-300 gpurun_out/c3_bench.json
