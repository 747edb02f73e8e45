#!/bin/bash
mkdir -p gpurun_out
SPECS="2:65536:2000:0:0 2:1024:10000:0:0 1:65536:2000:0:0 1:1:1000:0:0 3:65536:500:4:3 3:65536:500:8:3 16x8x16:65536:1500:2:3 16x8x16:4096:3000:0:0 3:65536:2000:8:0"
tools/ab_probe.sh "$SPECS" base > gpurun_out/r2s_ab_scan_batch.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2s_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2s_tests.log
cat gpurun_out/r2s_ab_scan_batch.log; tail -3 gpurun_out/r2s_tests.log
