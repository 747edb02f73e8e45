#!/bin/bash
# round-2 multi-GPU run: usage gpu_r2_multi.sh N   (run under gpurun --gpus N)
N=$1
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2m${N}_smi.txt 2>&1
timeout 900 python -m pytest tests/test_multi_device.py -m gpu -q --timeout 600 > gpurun_out/r2m${N}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2m${N}_tests.log
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
timeout 900 bash -c "$(declare -f run); N=$N; run 29511 --steps 3 --warmup 2" > gpurun_out/r2m${N}_bench.json 2> gpurun_out/r2m${N}_bench.err
echo "bench rc=$?" >> gpurun_out/r2m${N}_bench.err
timeout 600 bash -c "$(declare -f run); N=$N; run 29512 --scaling strong --steps 3 --warmup 2" > gpurun_out/r2m${N}_strong.json 2> gpurun_out/r2m${N}_strong.err
echo "strong rc=$?" >> gpurun_out/r2m${N}_strong.err
timeout 900 bash -c "$(declare -f run); N=$N; run 29513 --config 5 --chains 16384 --iterations 10000" > gpurun_out/r2m${N}_tempering.json 2> gpurun_out/r2m${N}_tempering.err
echo "tempering rc=$?" >> gpurun_out/r2m${N}_tempering.err
# the in-process path alone (one process, C ABI device list): plain C caller over all devices
tail -n 3 gpurun_out/r2m${N}_tests.log gpurun_out/r2m${N}_bench.err gpurun_out/r2m${N}_strong.err gpurun_out/r2m${N}_tempering.err
