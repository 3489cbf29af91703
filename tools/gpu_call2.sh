#!/bin/bash
# round-2 call 2: whole GPU suite with the new kernels, launch lists of the DCGAN step, the other configs' bench lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_baseline_configs.py -s > gpurun_out/c2_tests.log 2>&1
echo "suite exit $?" >> gpurun_out/c2_tests.log
timeout 900 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -q -s > gpurun_out/c2_tests_baseline.log 2>&1
echo "baseline-config tests exit $?" >> gpurun_out/c2_tests_baseline.log
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 600 $NCU --log-file gpurun_out/c2_launches_default.csv python bench.py --steps 2 --warmup 1 --no-gpu-reference --no-roofline --no-cpu-baseline > gpurun_out/c2_ncu_bench_default.log 2>&1
B200GAN_NARROW=1 timeout 600 $NCU --log-file gpurun_out/c2_launches_narrow.csv python bench.py --steps 2 --warmup 1 --no-gpu-reference --no-roofline --no-cpu-baseline > gpurun_out/c2_ncu_bench_narrow.log 2>&1
for cfg in wgan_gp pix2pix cyclegan; do
  timeout 600 python bench.py --config $cfg --no-cpu-baseline > gpurun_out/c2_bench_$cfg.json 2> gpurun_out/c2_bench_$cfg.err
done
tail -3 gpurun_out/c2_tests.log; tail -3 gpurun_out/c2_tests_baseline.log; cat gpurun_out/c2_bench_*.json | cut -c1-200
