#!/bin/bash
mkdir -p gpurun_out
SPECS=""
for n in 16 18 20 22 24 26 28; do c=$((n/2)); SPECS="$SPECS ${n}x${c}x${n}:65536:1500:0:3 ${n}x${c}x${n}:65536:1500:0:2"; done
tools/ab_probe.sh "$SPECS 2:65536:2000:0:3 2:65536:2000:0:2 2:1024:10000:0:3 2:1024:10000:0:2 1:65536:2000:0:3 1:65536:2000:0:2" base > gpurun_out/r2r_ab_threshold.log 2>&1
cat gpurun_out/r2r_ab_threshold.log
