#!/bin/bash
# development probe: which builds/launch shapes of the reference kernel terminate on this GPU
run() { echo "== $*"; timeout 45 python oracle/ref_runner.py "$@" 2>&1 | tail -2; echo "rc=$?"; }
run plain 1 64 100 1 0 1
run nb 1 64 100 64 0 1
run nb 3 1024 20 64 0 1
run nb 3 4096 50 64 1 1
run plain 3 4096 50 1 1 1
run nb 3 4096 50 32 0 1
run nb 2 1024 200 64 0 1
nvidia-smi --query-gpu=utilization.gpu,memory.used --format=csv
