#!/bin/bash
# round-2 call 1: validate the narrow kernels on hardware, time them, get the stock-torch DCGAN arm
mkdir -p gpurun_out
export B200GAN_NARROW=1
timeout 900 python -m pytest tests/test_gpu_y_narrow.py tests/test_gpu_ops.py tests/test_gpu_dcgan.py -m gpu -x -q > gpurun_out/c1_tests_narrow.log 2>&1
echo "narrow tests exit $?" >> gpurun_out/c1_tests_narrow.log
timeout 300 python tools/profile_kernels.py > gpurun_out/c1_kernels_narrow.log 2>&1
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c1_bench_narrow.json 2> gpurun_out/c1_bench_narrow.err
unset B200GAN_NARROW
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c1_bench_default.json 2> gpurun_out/c1_bench_default.err
timeout 300 python bench.py --impl stock --no-cpu-baseline > gpurun_out/c1_bench_stock.json 2> gpurun_out/c1_bench_stock.err
timeout 300 python bench.py --impl stock --no-cpu-baseline --no-graph > gpurun_out/c1_bench_stock_eager.json 2> gpurun_out/c1_bench_stock_eager.err
tail -3 gpurun_out/c1_tests_narrow.log; cat gpurun_out/c1_bench_*.json | cut -c1-300
