#!/bin/bash
# round-2 call 19: wgrad_tc with 128 x 256 output tiles (B200GAN_WG_NB256=1): parity, then speed; default build sanity
mkdir -p gpurun_out
B200GAN_WG_NB256=1 timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_models.py tests/test_gpu_baseline_configs.py -m gpu -q > gpurun_out/c19_tests_nb256.log 2>&1
echo "nb256 tests exit $?" >> gpurun_out/c19_tests_nb256.log
B200GAN_WG_NB256=1 timeout 200 python tools/profile_kernels.py > gpurun_out/c19_kernels_nb256.log 2>&1
B200GAN_WG_NB256=1 timeout 300 python bench.py --config cyclegan --no-cpu-baseline --no-gpu-reference --steps 5 --warmup 3 > gpurun_out/c19_bench_cyclegan_nb256.json 2> gpurun_out/c19_bench_cyclegan_nb256.err
B200GAN_WG_NB256=1 timeout 300 python bench.py --config pix2pix --no-cpu-baseline --no-gpu-reference > gpurun_out/c19_bench_pix2pix_nb256.json 2> gpurun_out/c19_bench_pix2pix_nb256.err
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/c19_tests_default.log 2>&1
echo "default tests exit $?" >> gpurun_out/c19_tests_default.log
tail -4 gpurun_out/c19_tests_nb256.log; grep -h "wgrad" gpurun_out/c19_kernels_nb256.log | grep "CG\|P2P"; for f in gpurun_out/c19_bench_*.json; do cut -c1-230 $f; done; tail -3 gpurun_out/c19_tests_default.log
