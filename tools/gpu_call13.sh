#!/bin/bash
# round-2 call 13: full GPU suite with the final kernels, planner sweep, the four BASELINE benches, small ncu captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c13_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/c13_tests.log
timeout 400 python tools/nb_sweep.py > gpurun_out/c13_nb_sweep.log 2>&1
timeout 400 python bench.py > gpurun_out/c13_bench.json 2> gpurun_out/c13_bench.err
timeout 200 python bench.py --config wgan_gp --no-cpu-baseline > gpurun_out/c13_bench_wgan_gp.json 2> gpurun_out/c13_bench_wgan_gp.err
timeout 300 python bench.py --config pix2pix --no-cpu-baseline > gpurun_out/c13_bench_pix2pix.json 2> gpurun_out/c13_bench_pix2pix.err
timeout 400 python bench.py --config cyclegan --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/c13_bench_cyclegan.json 2> gpurun_out/c13_bench_cyclegan.err
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 300 $NCU --log-file gpurun_out/c13_launches.csv python bench.py --steps 2 --warmup 1 --no-gpu-reference --no-roofline --no-cpu-baseline > gpurun_out/c13_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:'tail_' -c 3 -o gpurun_out/c13_tail_prof python bench.py --steps 1 --warmup 1 --no-gpu-reference --no-roofline --no-cpu-baseline --no-graph > gpurun_out/c13_ncu_tail.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:'fewk_|wgrad_reduce_tile|pack_multi' -c 6 -o gpurun_out/c13_edge_prof python bench.py --config pix2pix --steps 1 --warmup 1 --no-gpu-reference --no-roofline --no-cpu-baseline --no-graph > gpurun_out/c13_ncu_edge.log 2>&1
du -sh gpurun_out; tail -6 gpurun_out/c13_tests.log; cat gpurun_out/c13_nb_sweep.log | cut -c1-260; for f in gpurun_out/c13_bench*.json; do cut -c1-250 $f; done
