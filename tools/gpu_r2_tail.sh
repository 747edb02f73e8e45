#!/bin/bash
# how much does the partly filled last wave cost?  740 resident 4-warp blocks of 16 chains = 11840 chains per wave
mkdir -p gpurun_out
for c in 47360 59200 60000 62160 65536 68080 71040 82880; do
  python tools/prof_target.py 3 $c 2000 | tee -a gpurun_out/r2x_tail.log
done
