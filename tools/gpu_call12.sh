#!/bin/bash
# round-2 call 12: the DCGAN step with the grouped discriminator pass
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_dcgan.py tests/test_gpu_z_optimizer.py -m gpu -q > gpurun_out/c12_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/c12_tests.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c12_bench.json 2> gpurun_out/c12_bench.err
B200GAN_BATCH_D=0 timeout 300 python bench.py --no-cpu-baseline --no-gpu-reference --no-roofline > gpurun_out/c12_bench_separate.json 2> gpurun_out/c12_bench_separate.err
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 300 $NCU --log-file gpurun_out/c12_launches.csv python bench.py --steps 2 --warmup 1 --no-gpu-reference --no-roofline --no-cpu-baseline > gpurun_out/c12_ncu_bench.log 2>&1
tail -6 gpurun_out/c12_tests.log; cut -c1-300 gpurun_out/c12_bench.json; cut -c1-300 gpurun_out/c12_bench_separate.json; tail -2 gpurun_out/c12_bench.err
