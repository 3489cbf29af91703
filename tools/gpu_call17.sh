#!/bin/bash
mkdir -p gpurun_out
timeout 240 python tools/cyclegan_ops.py > gpurun_out/c17_cyclegan_ops.log 2>&1
cat gpurun_out/c17_cyclegan_ops.log
