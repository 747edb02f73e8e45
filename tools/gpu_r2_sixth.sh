#!/bin/bash
mkdir -p gpurun_out
python tools/latency_probe.py > gpurun_out/r2i_latency.log 2>&1
for r in 8 16 32; do
  timeout 600 python bench.py --config 5 --chains 16384 --iterations 10000 --rungs $r > gpurun_out/r2i_tempering_rungs$r.json 2> gpurun_out/r2i_tempering_rungs$r.err
done
timeout 600 python bench.py --config 5 --chains 16384 --iterations 10000 --rungs 16 --exchange-interval 25 > gpurun_out/r2i_tempering_rungs16_ex25.json 2> gpurun_out/r2i_tempering_rungs16_ex25.err
cat gpurun_out/r2i_latency.log
