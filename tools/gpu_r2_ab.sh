#!/bin/bash
mkdir -p gpurun_out
SPECS="3:65536:2000:8:0 4:65536:200:32:0 32x16x32:65536:1500:0:0 100x50x100:65536:400:0:0"
tools/ab_probe.sh "$SPECS" base pipe > gpurun_out/r2k_ab_pipe.log 2>&1
MH_LIB=$PWD/metropolis-hastings-gpgpu_b200/libKernel_pipe.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "memo or wild or default_mode or small_rooms" > gpurun_out/r2k_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2k_tests.log
cat gpurun_out/r2k_ab_pipe.log; tail -3 gpurun_out/r2k_tests.log
