#!/bin/bash
# development probe: A/B of experimental libKernel builds (MH_LIB)
for v in "$@"; do
  for l in 4 8; do
    echo -n "$v "; MH_LIB=$PWD/metropolis-hastings-gpgpu_b200/libKernel_$v.so python tests/prof_target.py 3 65536 200 $l
  done
  echo -n "$v "; MH_LIB=$PWD/metropolis-hastings-gpgpu_b200/libKernel_$v.so python tests/prof_target.py 4 16384 20 32
  echo -n "$v "; MH_LIB=$PWD/metropolis-hastings-gpgpu_b200/libKernel_$v.so python tests/prof_target.py 2 65536 500 2
done
