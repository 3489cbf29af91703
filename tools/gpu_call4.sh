#!/bin/bash
# round-2 call 4: shared-memory staged chain kernels (v2), slab-reduced weight gradient, test fixes
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nb_ops.py tests/test_gpu_chain.py -m gpu -q -x > gpurun_out/c4_tests_nb.log 2>&1
echo "nb tests exit $?" >> gpurun_out/c4_tests_nb.log
timeout 300 python tools/nb_bench.py > gpurun_out/c4_nb_bench.log 2>&1
B200GAN_NB_V1=1 timeout 300 python tools/nb_bench.py > gpurun_out/c4_nb_bench_v1.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_chain.py --deselect tests/test_gpu_nb_ops.py > gpurun_out/c4_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/c4_tests.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err
B200GAN_FUSE_CHAIN=0 timeout 600 python bench.py --no-cpu-baseline --no-gpu-reference --no-roofline > gpurun_out/c4_bench_nochain.json 2> gpurun_out/c4_bench_nochain.err
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 600 $NCU --log-file gpurun_out/c4_launches.csv python bench.py --steps 2 --warmup 1 --no-gpu-reference --no-roofline --no-cpu-baseline > gpurun_out/c4_ncu_bench.log 2>&1
timeout 300 python tools/profile_kernels.py > gpurun_out/c4_kernels.log 2>&1
tail -12 gpurun_out/c4_tests_nb.log; cat gpurun_out/c4_nb_bench.log; tail -8 gpurun_out/c4_tests.log; cut -c1-400 gpurun_out/c4_bench.json; cut -c1-200 gpurun_out/c4_bench_nochain.json
