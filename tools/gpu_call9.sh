#!/bin/bash
# round-2 call 9 (2 GPUs): the diagnostic NCCL test (local gradients + parameters), pix2pix golden with the yardstick bound
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_nccl_ddp.py -m gpu -q -s > gpurun_out/c9_ddp_tests.log 2>&1
echo "nccl ddp test exit $?" >> gpurun_out/c9_ddp_tests.log
B200GAN_DDP_OVERLAP=0 timeout 400 python -m pytest tests/test_gpu_nccl_ddp.py -m gpu -q -s > gpurun_out/c9_ddp_tests_nooverlap.log 2>&1
echo "nccl ddp test (no overlap) exit $?" >> gpurun_out/c9_ddp_tests_nooverlap.log
timeout 300 python -m pytest tests/test_gpu_models.py tests/test_gpu_z_optimizer.py tests/test_gpu_tail.py -m gpu -q > gpurun_out/c9_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/c9_tests.log
timeout 300 python bench.py --gpus 1 --no-cpu-baseline > gpurun_out/c9_bench_n1.json 2> gpurun_out/c9_bench_n1.err
grep -h "2-rank NCCL\|passed\|failed\|exit" gpurun_out/c9_ddp_tests.log gpurun_out/c9_ddp_tests_nooverlap.log gpurun_out/c9_tests.log | cut -c1-400; cut -c1-300 gpurun_out/c9_bench_n1.json
