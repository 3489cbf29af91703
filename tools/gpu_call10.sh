#!/bin/bash
# round-2 call 10: launch lists of the CycleGAN / DCGAN steps with the final kernels, ncu --set full of the new kernels
mkdir -p gpurun_out
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 300 $NCU --log-file gpurun_out/c10_launches.csv python bench.py --steps 2 --warmup 1 --no-gpu-reference --no-roofline --no-cpu-baseline > gpurun_out/c10_ncu_bench.log 2>&1
timeout 800 $NCU --log-file gpurun_out/c10_launches_cyclegan.csv python bench.py --config cyclegan --steps 1 --warmup 1 --no-gpu-reference --no-roofline --no-cpu-baseline --no-graph > gpurun_out/c10_ncu_cyclegan.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'fewk_|nbk_fprop2|nbk_dgrad2|nbk_wgrad_kernel|pack_multi|pad2d' -c 36 -o gpurun_out/c10_edge_prof python bench.py --config pix2pix --steps 1 --warmup 1 --no-gpu-reference --no-roofline --no-cpu-baseline --no-graph > gpurun_out/c10_ncu_edge.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'tail_' -c 6 -o gpurun_out/c10_tail_prof python bench.py --steps 1 --warmup 1 --no-gpu-reference --no-roofline --no-cpu-baseline --no-graph > gpurun_out/c10_ncu_tail.log 2>&1
ls -la gpurun_out/c10_*; tail -2 gpurun_out/c10_ncu_cyclegan.log | cut -c1-200
