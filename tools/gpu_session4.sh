#!/bin/bash
mkdir -p gpurun_out
for d in 1 2; do
  echo "== LCGP_DIAG=$d"
  CASES=128x1,128x32,256x1,1024x1,1024x8,2048x10,8064x1,8064x4 LCGP_DIAG=$d timeout 300 python tools/potrf_microbench.py
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/aux_kernels_bench.py
