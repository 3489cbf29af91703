#!/bin/bash
# round-2 call 18: conv_tc with cluster multicast of the weight boxes (B200GAN_TC_CLUSTER=1): parity, then speed
mkdir -p gpurun_out
B200GAN_TC_CLUSTER=1 timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x > gpurun_out/c18_tests_ops.log 2>&1
echo "ops tests exit $?" >> gpurun_out/c18_tests_ops.log
if grep -q "ops tests exit 0" gpurun_out/c18_tests_ops.log; then
  B200GAN_TC_CLUSTER=1 timeout 400 python -m pytest tests/test_gpu_models.py tests/test_gpu_baseline_configs.py tests/test_gpu_dcgan.py -m gpu -q > gpurun_out/c18_tests_models.log 2>&1
  echo "model tests exit $?" >> gpurun_out/c18_tests_models.log
  B200GAN_TC_CLUSTER=1 timeout 200 python tools/profile_kernels.py > gpurun_out/c18_kernels_cluster.log 2>&1
  timeout 200 python tools/profile_kernels.py > gpurun_out/c18_kernels_plain.log 2>&1
  B200GAN_TC_CLUSTER=1 timeout 300 python bench.py --config cyclegan --no-cpu-baseline --no-gpu-reference --steps 5 --warmup 3 > gpurun_out/c18_bench_cyclegan_cluster.json 2> gpurun_out/c18_bench_cyclegan_cluster.err
  B200GAN_TC_CLUSTER=1 timeout 300 python bench.py --config pix2pix --no-cpu-baseline --no-gpu-reference > gpurun_out/c18_bench_pix2pix_cluster.json 2> gpurun_out/c18_bench_pix2pix_cluster.err
  B200GAN_TC_CLUSTER=1 timeout 300 python bench.py --no-cpu-baseline --no-gpu-reference --no-roofline > gpurun_out/c18_bench_dcgan_cluster.json 2> gpurun_out/c18_bench_dcgan_cluster.err
fi
tail -5 gpurun_out/c18_tests_ops.log; tail -4 gpurun_out/c18_tests_models.log; grep -h "CG \|P2P " gpurun_out/c18_kernels_cluster.log gpurun_out/c18_kernels_plain.log; for f in gpurun_out/c18_bench_*.json; do cut -c1-230 $f; done
