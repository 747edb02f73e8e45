#!/bin/bash
# development probe: A/B of experimental libKernel builds (MH_LIB); "base" = the product build
for v in "$@"; do
  lib=$PWD/metropolis-hastings-gpgpu_b200/libKernel_$v.so; [ "$v" = base ] && lib=$PWD/metropolis-hastings-gpgpu_b200/libKernel.so
  echo -n "$v "; MH_LIB=$lib python tools/prof_target.py 3 65536 400 4 3 | tail -1
  echo -n "$v "; MH_LIB=$lib python tools/prof_target.py 3 65536 400 4 3 | tail -1
  echo -n "$v "; MH_LIB=$lib python tools/prof_target.py 2 65536 1000 2 3 | tail -1
  echo -n "$v "; MH_LIB=$lib python tools/prof_target.py 40,40,40 65536 400 0 3 | tail -1
done
