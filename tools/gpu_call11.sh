#!/bin/bash
# round-2 call 11: grouped discriminator pass (ops.bn_groups) -- parity, then the DCGAN step with it
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_chain.py tests/test_gpu_nb_ops.py tests/test_gpu_dcgan.py tests/test_gpu_n2.py -m gpu -q > gpurun_out/c11_tests_chain.log 2>&1
echo "chain tests exit $?" >> gpurun_out/c11_tests_chain.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c11_bench.json 2> gpurun_out/c11_bench.err
B200GAN_BATCH_D=0 timeout 300 python bench.py --no-cpu-baseline --no-gpu-reference --no-roofline > gpurun_out/c11_bench_separate.json 2> gpurun_out/c11_bench_separate.err
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 300 $NCU --log-file gpurun_out/c11_launches.csv python bench.py --steps 2 --warmup 1 --no-gpu-reference --no-roofline --no-cpu-baseline > gpurun_out/c11_ncu_bench.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_chain.py --deselect tests/test_gpu_nb_ops.py --deselect tests/test_gpu_dcgan.py --deselect tests/test_gpu_n2.py > gpurun_out/c11_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/c11_tests.log
tail -15 gpurun_out/c11_tests_chain.log; tail -6 gpurun_out/c11_tests.log; cut -c1-300 gpurun_out/c11_bench.json; cut -c1-300 gpurun_out/c11_bench_separate.json
