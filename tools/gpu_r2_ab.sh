#!/bin/bash
mkdir -p gpurun_out
SPECS="3:65536:2000:8:0 3:65536:2000:8:1 4:65536:200:32:0 32x16x32:65536:1500:0:0 100x50x100:65536:400:0:0 40x20x40:65536:1500:0:0 70x35x70:65536:800:0:0 3:1024:5000:0:0 3:4096:5000:0:0"
{
tools/ab_probe.sh "$SPECS" base
MH_DELTA_WARPS=6 tools/ab_probe.sh "3:65536:2000:8:0 4:65536:200:32:0 32x16x32:65536:1500:0:0 100x50x100:65536:400:0:0 40x20x40:65536:1500:0:0 70x35x70:65536:800:0:0" w6
MH_DELTA_WARPS=8 tools/ab_probe.sh "3:65536:2000:8:0 32x16x32:65536:1500:0:0 40x20x40:65536:1500:0:0 70x35x70:65536:800:0:0" w8
MH_DELTA_WARPS=4 tools/ab_probe.sh "4:65536:200:32:0 32x16x32:65536:1500:0:0 100x50x100:65536:400:0:0 40x20x40:65536:1500:0:0 70x35x70:65536:800:0:0" w4
} > gpurun_out/r2p_ab_shapes.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2p_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2p_tests.log
cat gpurun_out/r2p_ab_shapes.log; tail -3 gpurun_out/r2p_tests.log
