#!/bin/bash
# Round-2 GPU session D: batched (multi-emulator) evaluation + lock-step fits: tests, then config 5 on one GPU.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batched or emulators or lockstep or plan" > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2d_pytest.log
timeout 900 python bench.py --config cfg5_batch --emulators 64 --maxiter 30 --engine threads --no-cpu-baseline > gpurun_out/r2d_cfg5_threads_it30.json 2> gpurun_out/r2d_cfg5_threads_it30.err; echo "rc=$?"; cat gpurun_out/r2d_cfg5_threads_it30.json
timeout 900 python bench.py --config cfg5_batch --emulators 64 --maxiter 30 --engine lockstep --no-cpu-baseline > gpurun_out/r2d_cfg5_lockstep_it30.json 2> gpurun_out/r2d_cfg5_lockstep_it30.err; echo "rc=$?"; cat gpurun_out/r2d_cfg5_lockstep_it30.json; tail -3 gpurun_out/r2d_cfg5_lockstep_it30.err
timeout 900 python bench.py --config cfg5_batch --emulators 8 --maxiter 30 --engine lockstep --no-cpu-baseline > gpurun_out/r2d_cfg5_lockstep_e8_it30.json 2> gpurun_out/r2d_cfg5_lockstep_e8.err; echo "rc=$?"; cat gpurun_out/r2d_cfg5_lockstep_e8_it30.json
timeout 1500 python bench.py --config cfg5_batch --emulators 64 --engine lockstep > gpurun_out/r2d_cfg5_lockstep_full.json 2> gpurun_out/r2d_cfg5_lockstep_full.err; echo "rc=$?"; cat gpurun_out/r2d_cfg5_lockstep_full.json; tail -3 gpurun_out/r2d_cfg5_lockstep_full.err
