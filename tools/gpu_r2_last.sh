#!/bin/bash
# final-build captures of the plain scan kernel (proposal recipes batched) + a longer fuzz run with another seed
mkdir -p gpurun_out
for t in "r2k_scan_n16 2 65536 300 0 0" "r2l_scan_n50_g4 3 65536 100 4 3"; do
  set -- $t; name=$1; shift
  python tools/prof_target.py "$@" > gpurun_out/${name}_clean.log 2>&1 || continue
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mh_chain|mh_delta" -s 1 -c 1 -f -o gpurun_out/prof_${name} python tools/prof_target.py "$@" > gpurun_out/${name}_ncu.log 2>&1
  python tools/ncu_summary.py gpurun_out/prof_${name}.ncu-rep > gpurun_out/${name}_ncu_full.txt 2>> gpurun_out/${name}_ncu.log
  ncu -i gpurun_out/prof_${name}.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/${name}_source.csv.gz
  rm -f gpurun_out/prof_${name}.ncu-rep
done
timeout 400 python tools/fuzz_gpu.py 300 777 > gpurun_out/r2v_fuzz.log 2>&1
tail -3 gpurun_out/r2v_fuzz.log
head -3 gpurun_out/r2k_scan_n16_ncu_full.txt gpurun_out/r2l_scan_n50_g4_ncu_full.txt
