#!/bin/bash
# 2-GPU call: NCCL data-parallel correctness + the scaling bench at N = 1, 2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nccl_ddp.py -m gpu -q -s > gpurun_out/ddp_tests.log 2>&1
echo "nccl ddp test exit $?" >> gpurun_out/ddp_tests.log
timeout 300 python bench.py --gpus 1 --no-cpu-baseline --no-gpu-reference --no-roofline > gpurun_out/ddp_bench_n1.json 2> gpurun_out/ddp_bench_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 > gpurun_out/ddp_bench_n2.json 2> gpurun_out/ddp_bench_n2.err
B200GAN_DDP_OVERLAP=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 > gpurun_out/ddp_bench_n2_nooverlap.json 2> gpurun_out/ddp_bench_n2_nooverlap.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --config pix2pix --steps 10 > gpurun_out/ddp_bench_pix2pix_n2.json 2> gpurun_out/ddp_bench_pix2pix_n2.err
tail -5 gpurun_out/ddp_tests.log; for f in gpurun_out/ddp_bench_*.json; do echo $f; cut -c1-250 $f; done; tail -3 gpurun_out/ddp_bench_n2.err
