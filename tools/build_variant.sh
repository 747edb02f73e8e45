#!/bin/bash
# development probe: build an experimental libKernel_<name>.so with extra nvcc defines
#   tools/build_variant.sh w8s1 -DMH_WARPS_PER_BLOCK=8 -DMH_SYNC_ITER=1
set -e
name=$1; shift
here=$(cd "$(dirname "$0")/.." && pwd)
src=$here/metropolis-hastings-gpgpu_b200/csrc
out=$here/metropolis-hastings-gpgpu_b200/libKernel_$name.so
tmp=$(mktemp -d)
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr \
     "$@" -c $src/mh_kernels.cu -o $tmp/k.o
gcc -std=c11 -O2 -fPIC -fvisibility=hidden -D_GNU_SOURCE -c $src/kernel_wrapper.c -o $tmp/w.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out $tmp/k.o $tmp/w.o -lcudart -lm
rm -rf $tmp
echo built $out
