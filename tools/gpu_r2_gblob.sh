#!/bin/bash
# experiment: memo kernel's 4-warp build reading the problem blob from global memory (no per-block staging), 5 or 6 blocks per SM
mkdir -p gpurun_out
bash tools/ab_probe.sh "3:65536:2000:0:0 40x20x40:65536:1500:0:0 70x35x70:65536:800:0:0 32x16x32:65536:1500:0:0" base gb5 gb6 2>&1 | tee gpurun_out/r2y_ab_gblob.log
MH_LIB=$PWD/metropolis-hastings-gpgpu_b200/libKernel_gb6.so timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "memo or wild or identical" 2>&1 | tail -3 | tee -a gpurun_out/r2y_ab_gblob.log
