#!/bin/bash
# round-2 call 8 (2 GPUs): NCCL data-parallel correctness, N = 1 / 2 bench, and the tests touched since call 7
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_nccl_ddp.py -m gpu -q -s > gpurun_out/c8_ddp_tests.log 2>&1
echo "nccl ddp test exit $?" >> gpurun_out/c8_ddp_tests.log
timeout 500 python -m pytest tests/test_gpu_n2.py tests/test_gpu_chain.py tests/test_gpu_nb_ops.py tests/test_gpu_z_optimizer.py tests/test_gpu_models.py -m gpu -q > gpurun_out/c8_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/c8_tests.log
timeout 200 python tools/nb_bench.py > gpurun_out/c8_nb_bench.log 2>&1
timeout 300 python bench.py --gpus 1 --no-cpu-baseline --no-gpu-reference --no-roofline > gpurun_out/c8_bench_n1.json 2> gpurun_out/c8_bench_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 > gpurun_out/c8_bench_n2.json 2> gpurun_out/c8_bench_n2.err
B200GAN_DDP_OVERLAP=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --no-cpu-baseline --no-gpu-reference --no-roofline > gpurun_out/c8_bench_n2_nooverlap.json 2> gpurun_out/c8_bench_n2_nooverlap.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --config pix2pix --steps 10 --no-cpu-baseline --no-gpu-reference --no-roofline > gpurun_out/c8_bench_pix2pix_n2.json 2> gpurun_out/c8_bench_pix2pix_n2.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/c8_bench_ref_n2.json 2> gpurun_out/c8_bench_ref_n2.err
tail -5 gpurun_out/c8_ddp_tests.log; tail -6 gpurun_out/c8_tests.log; cat gpurun_out/c8_nb_bench.log | tail -7; for f in gpurun_out/c8_bench_*.json; do echo $f; cut -c1-260 $f; done; tail -3 gpurun_out/c8_bench_n2.err
