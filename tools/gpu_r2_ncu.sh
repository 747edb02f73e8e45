#!/bin/bash
mkdir -p gpurun_out
for t in "r2f_memo_n50_g8 3 65536 200 8 0"; do
  set -- $t; name=$1; shift
  python tools/prof_target.py "$@" > gpurun_out/${name}_clean.log 2>&1 || continue
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mh_chain|mh_delta" -s 1 -c 1 -f -o gpurun_out/prof_${name} python tools/prof_target.py "$@" > gpurun_out/${name}_ncu.log 2>&1
  python tools/ncu_summary.py gpurun_out/prof_${name}.ncu-rep > gpurun_out/${name}_ncu_full.txt 2>> gpurun_out/${name}_ncu.log
  ncu -i gpurun_out/prof_${name}.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/${name}_source.csv.gz
done
cat gpurun_out/r2f_memo_n50_g8_clean.log
