#!/bin/bash
mkdir -p gpurun_out
SPECS="3:65536:2000:8:0 3:65536:2000:8:1 4:65536:200:32:0 32x16x32:65536:1500:0:0 100x50x100:65536:400:0:0 3:1024:5000:0:0"
tools/ab_probe.sh "$SPECS" base > gpurun_out/r2n_ab_batch.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2n_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2n_tests.log
cat gpurun_out/r2n_ab_batch.log; tail -3 gpurun_out/r2n_tests.log
