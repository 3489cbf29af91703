#!/bin/bash
# round-2 call 15: final build -- full GPU suite, smoke, the four BASELINE benches, Pix2Pix / DCGAN launch lists
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c15_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/c15_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/c15_smoke.log 2>&1
timeout 400 python bench.py > gpurun_out/c15_bench.json 2> gpurun_out/c15_bench.err
timeout 200 python bench.py --config wgan_gp --no-cpu-baseline > gpurun_out/c15_bench_wgan_gp.json 2> gpurun_out/c15_bench_wgan_gp.err
timeout 300 python bench.py --config pix2pix --no-cpu-baseline > gpurun_out/c15_bench_pix2pix.json 2> gpurun_out/c15_bench_pix2pix.err
timeout 400 python bench.py --config cyclegan --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/c15_bench_cyclegan.json 2> gpurun_out/c15_bench_cyclegan.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c15_bench_reference.json 2> gpurun_out/c15_bench_reference.err
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 300 $NCU --log-file gpurun_out/c15_launches.csv python bench.py --steps 2 --warmup 1 --no-gpu-reference --no-roofline --no-cpu-baseline > gpurun_out/c15_ncu_bench.log 2>&1
timeout 420 $NCU --log-file gpurun_out/c15_launches_pix2pix.csv python bench.py --config pix2pix --steps 1 --warmup 1 --no-gpu-reference --no-roofline --no-cpu-baseline --no-graph > gpurun_out/c15_ncu_pix2pix.log 2>&1
tail -4 gpurun_out/c15_tests.log; tail -2 gpurun_out/c15_smoke.log; for f in gpurun_out/c15_bench*.json; do cut -c1-230 $f; done
