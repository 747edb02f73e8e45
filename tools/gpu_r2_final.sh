#!/bin/bash
# round-2 final validation on one GPU: the whole GPU suite, the bounds-checking build on the memo / delta / multi-device tests, smoke, bench
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2f_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2f_tests.log
if [ -f metropolis-hastings-gpgpu_b200/libKernel_dbg.so ]; then
MH_LIB=$PWD/metropolis-hastings-gpgpu_b200/libKernel_dbg.so timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_multi_device.py -m gpu -q --timeout 900 -k "memo or delta or wild or multi or chunk or tempering or ranking or ladder or frozen or trajector" > gpurun_out/r2f_tests_dbg_bounds.log 2>&1
echo "dbg tests rc=$?" >> gpurun_out/r2f_tests_dbg_bounds.log
fi
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2f_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
echo "bench rc=$?" >> gpurun_out/r2f_bench.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err
python tools/probe_lanes.py > gpurun_out/r2f_probe_lanes.log 2>&1
tail -n 3 gpurun_out/r2f_tests.log gpurun_out/r2f_tests_dbg_bounds.log gpurun_out/r2f_smoke.log gpurun_out/r2f_bench.err
