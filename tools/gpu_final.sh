#!/bin/bash
# Round-end style verification: GPU parity suite, smoke, default bench (developer tool).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; tail -3 gpurun_out/pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_1gpu_final.json 2> gpurun_out/bench_1gpu_final.err; tail -c 300 gpurun_out/bench_1gpu_final.json
timeout 600 python bench.py --config cfg5_batch 2>/dev/null | cut -c1-220
