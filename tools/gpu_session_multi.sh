#!/bin/bash
# usage: gpu_session_multi.sh N [fit]   (developer tool; outputs under gpurun_out/)
N=$1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/nccl_worker.py > gpurun_out/nccl_worker_$N.log 2>&1; tail -5 gpurun_out/nccl_worker_$N.log
FIT=""
if [ "$2" == "fit" ]; then FIT="--fit-config workload"; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 $FIT > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; tail -c 1500 gpurun_out/bench_${N}gpu.json
