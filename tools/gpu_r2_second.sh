#!/bin/bash
# round-2 second GPU run: the whole GPU test suite, parity table, ncu captures of the current kernels, launch list, e2e pin probe
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2b_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2b_tests.log
timeout 600 python tools/parity_errors.py 100000 > gpurun_out/r2b_parity.log 2>&1
cp profiles/parity_errors_r2.json gpurun_out/ 2>/dev/null
SPECS="3:65536:2000:8:0 3:65536:2000:4:3 3:65536:2000:8:3 2:65536:2000:0:0 2:1024:10000:0:0 1:65536:2000:0:0 1:1:1000:0:0 4:65536:200:32:0 24x12x24:65536:1500:0:0 24x12x24:65536:1500:4:2 20x10x20:65536:1500:0:0 20x10x20:65536:1500:4:2 20x10x20:65536:1500:2:2 28x14x28:65536:1500:0:0 28x14x28:65536:1500:4:3"
tools/ab_probe.sh "$SPECS" base > gpurun_out/r2b_ab.log 2>&1
# e2e at config-4 size with and without page-locking the result block
{
for pin in 0 1; do
  if [ $pin = 1 ]; then export MH_PIN_RESULT=1; else unset MH_PIN_RESULT; fi
  MH_TIMING=1 python tools/e2e_probe.py 4 262144 100 2>&1 | tail -4
  MH_TIMING=1 python tools/e2e_probe.py 3 65536 500 2>&1 | tail -4
done
unset MH_PIN_RESULT
} > gpurun_out/r2b_e2e_pin.log 2>&1
# ncu: each target first runs clean, then under ncu
for t in "r2a_memo_n50_g8 3 65536 200 8 0" "r2b_scan_n16 2 65536 300 0 0" "r2c_memo_n200_g32 4 16384 60 32 0" "r2d_scan_n50_g4 3 65536 100 4 3" "r2e_delta_n50_g8 3 65536 200 8 1"; do
  set -- $t; name=$1; shift
  python tools/prof_target.py "$@" > gpurun_out/${name}_clean.log 2>&1 || continue
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mh_chain|mh_delta" -s 1 -c 1 -f -o gpurun_out/prof_${name} python tools/prof_target.py "$@" > gpurun_out/${name}_ncu.log 2>&1
  # gpurun brings back at most 64 MiB: condense on the box, keep only the first capture's .ncu-rep
  python tools/ncu_summary.py gpurun_out/prof_${name}.ncu-rep > gpurun_out/${name}_ncu_full.txt 2>> gpurun_out/${name}_ncu.log
  ncu -i gpurun_out/prof_${name}.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/${name}_source.csv.gz
  [ "$name" = r2a_memo_n50_g8 ] || rm -f gpurun_out/prof_${name}.ncu-rep
done
python bench.py --steps 2 --warmup 3 --iterations 200 --no-cpu-baseline --no-ref-gpu > gpurun_out/r2b_bench_iters200.json 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench_iters200.csv python bench.py --steps 2 --warmup 3 --iterations 200 --no-cpu-baseline --no-ref-gpu > gpurun_out/r2b_ncu_bench.log 2>&1
tail -n 5 gpurun_out/r2b_tests.log gpurun_out/r2b_ab.log
