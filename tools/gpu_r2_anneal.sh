#!/bin/bash
# the annealing sub-record of the default bench line and the five-sampler time-to-target run
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --config 5 --chains 16384 --iterations 10000 > gpurun_out/r2w_tempering.json 2> gpurun_out/r2w_tempering.err; echo "tempering rc=$?"
timeout 600 python bench.py --config 5 --chains 16384 --iterations 10000 --rungs 32 > gpurun_out/r2w_tempering_rungs32.json 2> gpurun_out/r2w_tempering_rungs32.err; echo "tempering32 rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2w_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], json.dumps(d.get('config3_annealing'))[:900])
for f in ('r2w_tempering','r2w_tempering_rungs32'):
    t=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    print(f, json.dumps(t['tempering']['summary']))
P
