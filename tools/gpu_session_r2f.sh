#!/bin/bash
# Round-2 GPU session F (2 GPUs): NCCL-sharded parity test, sharded bench smoke, config 5 on 2 ranks; full suite.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nccl_sharded.py -m gpu -x -q -s > gpurun_out/r2f_pytest_nccl.log 2>&1; echo "nccl pytest rc=$?"; grep -E "NCCL_SHARDED|\[|passed|failed" gpurun_out/r2f_pytest_nccl.log | tail -12
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2f_pytest_gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-fit > gpurun_out/r2f_bench_2gpu.json 2> gpurun_out/r2f_bench_2gpu.err; echo "bench2 rc=$?"; cut -c1-900 gpurun_out/r2f_bench_2gpu.json; tail -3 gpurun_out/r2f_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config cfg5_batch --emulators 16 --maxiter 40 > gpurun_out/r2f_cfg5_2gpu.json 2> gpurun_out/r2f_cfg5_2gpu.err; echo "cfg5 rc=$?"; cat gpurun_out/r2f_cfg5_2gpu.json; tail -3 gpurun_out/r2f_cfg5_2gpu.err
