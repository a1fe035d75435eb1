#!/bin/bash
# Round-2 GPU session B: first run of the persistent left-looking Cholesky kernel and the blocked diagonal-block kernel.
mkdir -p gpurun_out
echo "== stage tests (persistent + launch chain with diag v2)"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "potrf or stages" > gpurun_out/r2b_pytest_potrf.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r2b_pytest_potrf.log
echo "== microbench: persistent kernel"
timeout 300 python tools/potrf_microbench.py > gpurun_out/r2b_microbench_pll.log 2>&1; echo "rc=$?"; cat gpurun_out/r2b_microbench_pll.log
echo "== microbench: launch chain + diag v2"
LCGP_POTRF=panels timeout 300 python tools/potrf_microbench.py > gpurun_out/r2b_microbench_panels_v2.log 2>&1; echo "rc=$?"; cat gpurun_out/r2b_microbench_panels_v2.log
echo "== full gpu suite"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2b_pytest_gpu.log
timeout 200 python tools/cfg5_latency.py cfg5_one > gpurun_out/r2b_latency.log 2>&1
timeout 200 python tools/cfg5_latency.py cfg3_rep >> gpurun_out/r2b_latency.log 2>&1; cat gpurun_out/r2b_latency.log
