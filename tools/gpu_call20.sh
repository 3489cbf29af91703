#!/bin/bash
# round-2 call 20: final build (128x256 weight-gradient tiles on) -- full GPU suite, smoke, the four BASELINE benches, Pix2Pix / DCGAN launch lists
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c20_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/c20_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/c20_smoke.log 2>&1
timeout 400 python bench.py > gpurun_out/c20_bench.json 2> gpurun_out/c20_bench.err
timeout 200 python bench.py --config wgan_gp --no-cpu-baseline > gpurun_out/c20_bench_wgan_gp.json 2> gpurun_out/c20_bench_wgan_gp.err
timeout 300 python bench.py --config pix2pix --no-cpu-baseline > gpurun_out/c20_bench_pix2pix.json 2> gpurun_out/c20_bench_pix2pix.err
timeout 400 python bench.py --config cyclegan --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/c20_bench_cyclegan.json 2> gpurun_out/c20_bench_cyclegan.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c20_bench_reference.json 2> gpurun_out/c20_bench_reference.err
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
tail -4 gpurun_out/c20_tests.log; tail -2 gpurun_out/c20_smoke.log; for f in gpurun_out/c20_bench*.json; do cut -c1-230 $f; done
