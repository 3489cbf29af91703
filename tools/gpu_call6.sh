#!/bin/bash
# round-2 call 6: few-output-channel kernels (fewk.cu) -- parity, then the pix2pix / cyclegan steps with them
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fewk.py -m gpu -q > gpurun_out/c6_tests_fewk.log 2>&1
echo "fewk tests exit $?" >> gpurun_out/c6_tests_fewk.log
timeout 900 python -m pytest tests/test_gpu_baseline_configs.py tests/test_gpu_models.py tests/test_gpu_ops.py -m gpu -q > gpurun_out/c6_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/c6_tests.log
timeout 600 python bench.py --config pix2pix --no-cpu-baseline > gpurun_out/c6_bench_pix2pix.json 2> gpurun_out/c6_bench_pix2pix.err
timeout 900 python bench.py --config cyclegan --no-cpu-baseline --steps 5 --warmup 2 > gpurun_out/c6_bench_cyclegan.json 2> gpurun_out/c6_bench_cyclegan.err
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 900 $NCU --log-file gpurun_out/c6_launches_pix2pix.csv python bench.py --config pix2pix --steps 1 --warmup 1 --no-gpu-reference --no-roofline --no-cpu-baseline --no-graph > gpurun_out/c6_ncu_pix2pix.log 2>&1
timeout 1200 $NCU --log-file gpurun_out/c6_launches_cyclegan.csv python bench.py --config cyclegan --steps 1 --warmup 1 --no-gpu-reference --no-roofline --no-cpu-baseline --no-graph > gpurun_out/c6_ncu_cyclegan.log 2>&1
tail -15 gpurun_out/c6_tests_fewk.log; tail -8 gpurun_out/c6_tests.log; cut -c1-300 gpurun_out/c6_bench_pix2pix.json; cut -c1-300 gpurun_out/c6_bench_cyclegan.json
