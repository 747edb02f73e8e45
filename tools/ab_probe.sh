#!/bin/bash
# development probe: A/B of experimental libKernel builds (MH_LIB); "base" = the product build
#   tools/ab_probe.sh "<run specs>" base w8s1 ...
specs=$1; shift
for v in "$@"; do
  lib=$PWD/metropolis-hastings-gpgpu_b200/libKernel_$v.so; [ -f "$lib" ] || lib=$PWD/metropolis-hastings-gpgpu_b200/libKernel.so
  MH_LIB=$lib python tools/ab_matrix.py $v $specs 2>&1 | cut -c 1-160
done
