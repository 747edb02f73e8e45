#!/bin/bash
# round-2 first GPU shake-down: new tests, parity error table, A/B of the memo-kernel launch shapes, a short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 tests/test_propose_parity.py tests/test_multi_device.py > gpurun_out/r2a_tests_new.log 2>&1
echo "new tests rc=$?" >> gpurun_out/r2a_tests_new.log
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 tests/test_gpu_parity.py tests/test_ks_parity.py > gpurun_out/r2a_tests_old.log 2>&1
echo "old tests rc=$?" >> gpurun_out/r2a_tests_old.log
timeout 600 python tools/parity_errors.py 100000 > gpurun_out/r2a_parity.log 2>&1
cp profiles/parity_errors_r2.json gpurun_out/ 2>/dev/null
SPECS="3:65536:2000:8:0 3:65536:2000:8:3 3:65536:2000:8:1 3:65536:2000:4:0 3:65536:2000:16:0 4:65536:200:32:0 4:65536:200:32:1 4:65536:100:32:3 2:65536:2000:0:0 2:65536:2000:2:2 2:1024:10000:0:0 1:65536:2000:0:0 24x12x24:65536:1500:0:0 24x12x24:65536:1500:4:2 32x16x32:65536:1500:0:0 32x16x32:65536:1500:4:3"
{
tools/ab_probe.sh "$SPECS" base
MH_DELTA_WARPS=6 MH_DELTA_REG_WARPS=18 tools/ab_probe.sh "3:65536:2000:8:0 3:65536:2000:8:1 4:65536:200:32:0" d3x6
MH_DELTA_WARPS=4 MH_DELTA_REG_WARPS=20 tools/ab_probe.sh "3:65536:2000:8:0 3:65536:2000:8:1 4:65536:200:32:0" d5x4
MH_DELTA_WARPS=4 tools/ab_probe.sh "3:65536:2000:8:0 3:65536:2000:8:1 4:65536:200:32:0" d4x4
tools/ab_probe.sh "2:65536:2000:0:0 2:1024:10000:0:0 1:65536:2000:0:0 3:65536:500:4:3 24x12x24:65536:1500:0:0" c4
} > gpurun_out/r2a_ab.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?" >> gpurun_out/r2a_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err
tail -5 gpurun_out/r2a_tests_new.log gpurun_out/r2a_tests_old.log gpurun_out/r2a_ab.log
