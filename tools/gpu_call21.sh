#!/bin/bash
# round-2 call 21: ncu --set full of the generator-tail backward kernels (final build) + DCGAN launch list
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none -k regex:'tail_bwd' -c 2 -o gpurun_out/c21_tail_bwd_prof python bench.py --steps 1 --warmup 1 --no-gpu-reference --no-roofline --no-cpu-baseline --no-graph > gpurun_out/c21_ncu_tail.log 2>&1
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 300 $NCU --log-file gpurun_out/c21_launches.csv python bench.py --steps 2 --warmup 1 --no-gpu-reference --no-roofline --no-cpu-baseline > gpurun_out/c21_ncu_bench.log 2>&1
ls -la gpurun_out/c21_*
