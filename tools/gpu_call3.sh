#!/bin/bash
# round-2 call 3: fused D chain, rewritten tail kernels, norm fast paths, one-kernel critic step, optimizer-driven packs
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_chain.py -m gpu -q > gpurun_out/c3_tests_chain.log 2>&1
echo "chain tests exit $?" >> gpurun_out/c3_tests_chain.log
timeout 1500 python -m pytest tests -m gpu -q -s --deselect tests/test_gpu_chain.py > gpurun_out/c3_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/c3_tests.log
if ! grep -q "tests exit 0" gpurun_out/c3_tests.log; then
  SUB="tests/test_gpu_ops.py tests/test_gpu_models.py tests/test_gpu_dcgan.py tests/test_gpu_tail.py"
  B200GAN_TC_BN256=0 B200GAN_WG_PIX=32 B200GAN_TC_KSPLIT=0 timeout 600 python -m pytest $SUB -m gpu -q > gpurun_out/c3_tests_tc_switches_off.log 2>&1
  echo "tc-switches-off exit $?" >> gpurun_out/c3_tests_tc_switches_off.log
  B200GAN_NORM_FAST=0 timeout 600 python -m pytest $SUB -m gpu -q > gpurun_out/c3_tests_norm_fast_off.log 2>&1
  echo "norm-fast-off exit $?" >> gpurun_out/c3_tests_norm_fast_off.log
  B200GAN_FUSE_CHAIN=0 B200GAN_FUSE_TAIL=0 timeout 600 python -m pytest $SUB -m gpu -q > gpurun_out/c3_tests_fusions_off.log 2>&1
  echo "fusions-off exit $?" >> gpurun_out/c3_tests_fusions_off.log
fi
timeout 300 python tools/profile_kernels.py > gpurun_out/c3_kernels.log 2>&1
B200GAN_TC_BN256=0 B200GAN_WG_PIX=32 timeout 300 python tools/profile_kernels.py > gpurun_out/c3_kernels_switches_off.log 2>&1
B200GAN_WG_PIX128=64 timeout 300 python tools/profile_kernels.py > gpurun_out/c3_kernels_wgpix128.log 2>&1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err
timeout 300 python bench.py --config wgan_gp --no-cpu-baseline > gpurun_out/c3_bench_wgan_gp.json 2> gpurun_out/c3_bench_wgan_gp.err
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 600 $NCU --log-file gpurun_out/c3_launches.csv python bench.py --steps 2 --warmup 1 --no-gpu-reference --no-roofline --no-cpu-baseline > gpurun_out/c3_ncu_bench.log 2>&1
timeout 900 $NCU --log-file gpurun_out/c3_launches_pix2pix.csv python bench.py --config pix2pix --steps 1 --warmup 1 --no-gpu-reference --no-roofline --no-cpu-baseline > gpurun_out/c3_ncu_pix2pix.log 2>&1
timeout 900 $NCU --log-file gpurun_out/c3_launches_cyclegan.csv python bench.py --config cyclegan --steps 1 --warmup 1 --no-gpu-reference --no-roofline --no-cpu-baseline > gpurun_out/c3_ncu_cyclegan.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tail_ -c 6 -o gpurun_out/c3_tail_prof python bench.py --steps 2 --warmup 1 --no-gpu-reference --no-cpu-baseline > gpurun_out/c3_ncu_tail.log 2>&1
tail -15 gpurun_out/c3_tests_chain.log; tail -8 gpurun_out/c3_tests.log; cut -c1-300 gpurun_out/c3_bench.json
