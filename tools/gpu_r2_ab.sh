#!/bin/bash
mkdir -p gpurun_out
SPECS="50x25x50:65536:2000:8:0 50x25x50:65536:2000:4:0 50x25x20:65536:2000:8:0 50x25x20:65536:2000:4:0 50x25x8:65536:2000:8:0 50x25x8:65536:2000:4:0 50x10x8:65536:2000:8:0 50x10x8:65536:2000:4:0"
tools/ab_probe.sh "$SPECS" base > gpurun_out/r2g_ab_g4.log 2>&1
MH_DELTA_WARPS=6 tools/ab_probe.sh "50x25x20:65536:2000:4:0 50x25x8:65536:2000:4:0" base >> gpurun_out/r2g_ab_g4.log 2>&1
cat gpurun_out/r2g_ab_g4.log
