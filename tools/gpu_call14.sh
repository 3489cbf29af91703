#!/bin/bash
# round-2 call 14: folded x2-upsample fewk kernels -- parity and the Pix2Pix step
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_fewk.py tests/test_gpu_models.py tests/test_gpu_baseline_configs.py -m gpu -q > gpurun_out/c14_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/c14_tests.log
timeout 300 python bench.py --config pix2pix --no-cpu-baseline > gpurun_out/c14_bench_pix2pix.json 2> gpurun_out/c14_bench_pix2pix.err
B200GAN_FEWK_FOLD=0 timeout 300 python bench.py --config pix2pix --no-cpu-baseline --no-gpu-reference > gpurun_out/c14_bench_pix2pix_nofold.json 2> gpurun_out/c14_bench_pix2pix_nofold.err
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 420 $NCU --log-file gpurun_out/c14_launches_pix2pix.csv python bench.py --config pix2pix --steps 1 --warmup 1 --no-gpu-reference --no-roofline --no-cpu-baseline --no-graph > gpurun_out/c14_ncu_pix2pix.log 2>&1
tail -6 gpurun_out/c14_tests.log; cut -c1-260 gpurun_out/c14_bench_pix2pix.json; cut -c1-260 gpurun_out/c14_bench_pix2pix_nofold.json
