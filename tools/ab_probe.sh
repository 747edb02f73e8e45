#!/bin/bash
# development probe: A/B of experimental libKernel builds (MH_LIB); "base" = the product build
for v in "$@"; do
  lib=$PWD/metropolis-hastings-gpgpu_b200/libKernel_$v.so; [ "$v" = base ] && lib=$PWD/metropolis-hastings-gpgpu_b200/libKernel.so
  for m in 1 2; do for l in 8 16; do echo -n "$v "; MH_LIB=$lib python tools/prof_target.py 3 65536 400 $l $m 2>&1 | tail -1 | cut -c 1-150; done; done
  echo -n "$v "; MH_LIB=$lib python tools/prof_target.py 4 16384 60 32 1 2>&1 | tail -1 | cut -c 1-150
  echo -n "$v "; MH_LIB=$lib python tools/prof_target.py 2 65536 1000 2 1 2>&1 | tail -1 | cut -c 1-150
  echo -n "$v "; MH_LIB=$lib python tools/prof_target.py 2 65536 1000 4 1 2>&1 | tail -1 | cut -c 1-150
done
