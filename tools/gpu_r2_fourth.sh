#!/bin/bash
# round-2 fourth GPU run: tests of the chunked one-shot path and multi-device distinct top-k, e2e with / without chunking, launch list of the headline step
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_multi_device.py -m gpu -q --timeout 600 > gpurun_out/r2d_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2d_tests.log
{
for ch in 1 ""; do
  if [ -n "$ch" ]; then export MH_CHUNKS=$ch; else unset MH_CHUNKS; fi
  echo "== MH_CHUNKS=${MH_CHUNKS:-default}"
  MH_TIMING=1 python tools/e2e_probe.py 4 262144 1000 2>&1 | tail -5
  MH_TIMING=1 python tools/e2e_probe.py 4 32768 1000 2>&1 | tail -3
  MH_TIMING=1 python tools/e2e_probe.py 3 65536 2000 2>&1 | tail -3
done
unset MH_CHUNKS
} > gpurun_out/r2d_e2e_chunks.log 2>&1
timeout 600 python bench.py --scaling strong --steps 3 --warmup 2 > gpurun_out/r2d_strong_n1.json 2> gpurun_out/r2d_strong_n1.err
python bench.py --steps 2 --warmup 3 --iterations 200 --no-extras --no-cpu-baseline --no-ref-gpu > gpurun_out/r2d_bench_iters200.json 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench_iters200.csv python bench.py --steps 2 --warmup 3 --iterations 200 --no-extras --no-cpu-baseline --no-ref-gpu > gpurun_out/r2d_ncu_bench.log 2>&1
tail -n 4 gpurun_out/r2d_tests.log gpurun_out/r2d_e2e_chunks.log
