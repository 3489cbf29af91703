#!/bin/bash
# round-2 call 22: tail backward with group-shared gradient loads -- parity and the DCGAN step
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tail.py tests/test_gpu_dcgan.py -m gpu -q > gpurun_out/c22_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/c22_tests.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c22_bench.json 2> gpurun_out/c22_bench.err
tail -4 gpurun_out/c22_tests.log; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c22_bench.json').read().strip().splitlines()[-1])
print('ms', d['ms_per_step'], 'speedup', d.get('speedup_vs_gpu_reference'))
for g in d['roofline']['groups']:
    if 'tail' in g['kernel']: print(g['kernel'][:60], round(g['ms']*1000,1), 'us frac', round(g['frac'],3))
PY
