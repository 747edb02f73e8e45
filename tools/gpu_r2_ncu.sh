#!/bin/bash
mkdir -p gpurun_out
for t in "r2g_memo_n50_g8 3 65536 200 8 0" "r2h_memo_n200_g32 4 16384 60 32 0" "r2i_scan_n50_g4 3 65536 100 4 3" "r2j_scan_n16 2 65536 300 0 0"; do
  set -- $t; name=$1; shift
  python tools/prof_target.py "$@" > gpurun_out/${name}_clean.log 2>&1 || continue
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mh_chain|mh_delta" -s 1 -c 1 -f -o gpurun_out/prof_${name} python tools/prof_target.py "$@" > gpurun_out/${name}_ncu.log 2>&1
  python tools/ncu_summary.py gpurun_out/prof_${name}.ncu-rep > gpurun_out/${name}_ncu_full.txt 2>> gpurun_out/${name}_ncu.log
  ncu -i gpurun_out/prof_${name}.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/${name}_source.csv.gz
  [ "$name" = r2g_memo_n50_g8 ] || rm -f gpurun_out/prof_${name}.ncu-rep
done
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err
echo "bench rc=$?" >> gpurun_out/r2q_bench.err
python bench.py --steps 2 --warmup 3 --iterations 200 --no-extras --no-cpu-baseline --no-ref-gpu > gpurun_out/r2q_bench_iters200.json 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench_iters200.csv python bench.py --steps 2 --warmup 3 --iterations 200 --no-extras --no-cpu-baseline --no-ref-gpu > gpurun_out/r2q_ncu_bench.log 2>&1
timeout 300 python tools/parity_errors.py 100000 > gpurun_out/r2q_parity.log 2>&1
cp profiles/parity_errors_r2.json gpurun_out/
tail -2 gpurun_out/r2q_bench.err
