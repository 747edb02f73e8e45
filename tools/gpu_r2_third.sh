#!/bin/bash
# round-2 third GPU run: informational KS vs the rebuilt reference, compute-sanitizer attempt, config-5 harness at N=1, bench
mkdir -p gpurun_out
timeout 900 python tools/ks_vs_reference.py > gpurun_out/r2c_ks_vs_ref.log 2>&1
cp profiles/r2_ks_vs_reference.json gpurun_out/ 2>/dev/null
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_target.py > gpurun_out/r2c_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/r2c_sanitizer_$tool.log
done
tools/ab_probe.sh "3:65536:2000:8:0 4:65536:200:32:0 3:65536:2000:8:1" base row2 scan2 clr2 rs2 > gpurun_out/r2c_ab_unroll.log 2>&1
timeout 600 python bench.py --config 5 --chains 16384 --iterations 10000 > gpurun_out/r2c_tempering_n1.json 2> gpurun_out/r2c_tempering_n1.err
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
echo "bench rc=$?" >> gpurun_out/r2c_bench.err
tail -n 3 gpurun_out/r2c_ks_vs_ref.log gpurun_out/r2c_sanitizer_*.log gpurun_out/r2c_bench.err
