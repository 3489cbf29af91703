#!/bin/bash
# round-2 call 7: register-tiled chain kernels (PT pixels per thread), table-driven weight gradient, edge-layer routes
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nb_ops.py tests/test_gpu_chain.py tests/test_gpu_fewk.py -m gpu -q > gpurun_out/c7_tests_nb.log 2>&1
echo "nb tests exit $?" >> gpurun_out/c7_tests_nb.log
timeout 300 python tools/nb_bench.py > gpurun_out/c7_nb_bench.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_chain.py --deselect tests/test_gpu_nb_ops.py --deselect tests/test_gpu_fewk.py > gpurun_out/c7_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/c7_tests.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err
timeout 600 python bench.py --config pix2pix --no-cpu-baseline > gpurun_out/c7_bench_pix2pix.json 2> gpurun_out/c7_bench_pix2pix.err
timeout 600 python bench.py --config cyclegan --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/c7_bench_cyclegan.json 2> gpurun_out/c7_bench_cyclegan.err
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 300 $NCU --log-file gpurun_out/c7_launches.csv python bench.py --steps 2 --warmup 1 --no-gpu-reference --no-roofline --no-cpu-baseline > gpurun_out/c7_ncu_bench.log 2>&1
timeout 420 $NCU --log-file gpurun_out/c7_launches_pix2pix.csv python bench.py --config pix2pix --steps 1 --warmup 1 --no-gpu-reference --no-roofline --no-cpu-baseline --no-graph > gpurun_out/c7_ncu_pix2pix.log 2>&1
tail -12 gpurun_out/c7_tests_nb.log; cat gpurun_out/c7_nb_bench.log; tail -8 gpurun_out/c7_tests.log; cut -c1-300 gpurun_out/c7_bench.json; cut -c1-300 gpurun_out/c7_bench_pix2pix.json; cut -c1-300 gpurun_out/c7_bench_cyclegan.json
