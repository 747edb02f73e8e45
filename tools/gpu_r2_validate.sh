#!/bin/bash
# round-end sequence on a fresh box: GPU suite, smoke, reference arm, bench
mkdir -p gpurun_out
( time python -m pytest tests/ -x -q -m gpu ) > gpurun_out/r2z_tests.log 2>&1; tail -4 gpurun_out/r2z_tests.log
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2z_smoke.log 2>&1; tail -5 gpurun_out/r2z_smoke.log
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2z_ref.log 2>&1; tail -4 gpurun_out/r2z_ref.log | cut -c 1-300
( time python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2z_bench.log 2>&1; tail -5 gpurun_out/r2z_bench.log | cut -c 1-600
