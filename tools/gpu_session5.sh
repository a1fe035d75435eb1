#!/bin/bash
mkdir -p gpurun_out
for d in 1 4 2 3; do
  echo "== LCGP_DIAG=$d"
  CASES=128x1,128x32,1024x1,1024x8,2048x10,8064x4 LCGP_DIAG=$d timeout 120 python tools/potrf_microbench.py
done
