#!/bin/bash
# Round-2 GPU session A: full GPU test suite (incl. the new named-shape parity tests), default bench line,
# isolated Cholesky timings and small-problem latencies (baseline numbers of this round before the kernel work).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2a_pytest_gpu.log
tail -5 gpurun_out/r2a_pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench_default.json 2> gpurun_out/r2a_bench_default.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2a_bench_default.json
timeout 300 python tools/potrf_microbench.py > gpurun_out/r2a_potrf_microbench.log 2>&1; cat gpurun_out/r2a_potrf_microbench.log
timeout 200 python tools/cfg5_latency.py cfg5_one > gpurun_out/r2a_latency.log 2>&1
timeout 200 python tools/cfg5_latency.py cfg3_rep >> gpurun_out/r2a_latency.log 2>&1; cat gpurun_out/r2a_latency.log
