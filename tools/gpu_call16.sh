#!/bin/bash
# round-2 call 16 (2 GPUs): NCCL data-parallel test with the final build, N = 1 / 2 benches, reference arm under torchrun
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_nccl_ddp.py -m gpu -q -s > gpurun_out/c16_ddp_tests.log 2>&1
echo "nccl ddp test exit $?" >> gpurun_out/c16_ddp_tests.log
timeout 300 python bench.py --gpus 1 --no-cpu-baseline --no-gpu-reference --no-roofline > gpurun_out/c16_bench_n1.json 2> gpurun_out/c16_bench_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 > gpurun_out/c16_bench_n2.json 2> gpurun_out/c16_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --config pix2pix --steps 10 --no-cpu-baseline --no-gpu-reference --no-roofline > gpurun_out/c16_bench_pix2pix_n2.json 2> gpurun_out/c16_bench_pix2pix_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29615 bench.py --gpus 2 --config cyclegan --steps 4 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-roofline > gpurun_out/c16_bench_cyclegan_n2.json 2> gpurun_out/c16_bench_cyclegan_n2.err
grep -h "2-rank NCCL\|passed\|failed\|exit" gpurun_out/c16_ddp_tests.log | cut -c1-400; for f in gpurun_out/c16_bench_*.json; do echo $f; grep -h "impl" $f | cut -c1-230; done; tail -2 gpurun_out/c16_bench_n2.err
