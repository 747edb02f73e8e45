#!/bin/bash
# first GPU shake-down: tests, then a quick throughput probe per lane width
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 600 python tools/probe_lanes.py > gpurun_out/probe_lanes.log 2>&1
cat gpurun_out/probe_lanes.log
