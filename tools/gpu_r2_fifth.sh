#!/bin/bash
# round-2 fifth GPU run: e2e with / without chunking (no MH_TIMING: its synchronise defeats the overlap), then the phases once
mkdir -p gpurun_out
{
for ch in 1 ""; do
  if [ -n "$ch" ]; then export MH_CHUNKS=$ch; else unset MH_CHUNKS; fi
  echo "== MH_CHUNKS=${MH_CHUNKS:-default}"
  python tools/e2e_probe.py 4 262144 1000 2>&1 | tail -5
  python tools/e2e_probe.py 4 32768 1000 2>&1 | tail -5
  python tools/e2e_probe.py 3 65536 2000 2>&1 | tail -5
done
unset MH_CHUNKS
echo "== phases (MH_TIMING=1, default chunks)"
MH_TIMING=1 python tools/e2e_probe.py 4 262144 1000 2>&1 | tail -9
MH_TIMING=1 python tools/e2e_probe.py 3 65536 2000 2>&1 | tail -9
} > gpurun_out/r2e_e2e_chunks.log 2>&1
timeout 900 python -m pytest tests/test_multi_device.py tests/test_gpu_parity.py -m gpu -q --timeout 600 -x > gpurun_out/r2e_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2e_tests.log
timeout 600 python bench.py --scaling strong --steps 3 --warmup 2 > gpurun_out/r2e_strong_n1.json 2> gpurun_out/r2e_strong_n1.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
cat gpurun_out/r2e_e2e_chunks.log; tail -n 3 gpurun_out/r2e_tests.log
