#!/bin/bash
# Developer GPU session: parity suite, Cholesky tuning sweep, benches (outputs under gpurun_out/).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_s2.log 2>&1; tail -3 gpurun_out/pytest_gpu_s2.log
{
for cfg in "8 0 4" "8 16 4" "8 24 4" "8 16 2" "4 0 4" "16 16 4" "16 24 8"; do
  set -- $cfg
  echo "== PANEL_W=$1 TAIL_N=$2 TAIL_W=$3"
  CASES=8064x1,8064x4,8064x8,4096x4,2048x10 LCGP_PANEL_W=$1 LCGP_TAIL_N=$2 LCGP_TAIL_W=$3 timeout 300 python tools/potrf_microbench.py
done
} > gpurun_out/potrf_sweep.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_1gpu_s2.json 2> gpurun_out/bench_1gpu_s2.err; tail -c 600 gpurun_out/bench_1gpu_s2.json
timeout 600 python bench.py --config cfg5_batch > gpurun_out/bench_cfg5_s2.json 2> gpurun_out/bench_cfg5_s2.err; tail -c 400 gpurun_out/bench_cfg5_s2.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_s2.log 2>&1; tail -1 gpurun_out/smoke_s2.log
