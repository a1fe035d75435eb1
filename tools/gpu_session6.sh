#!/bin/bash
for m in 0 1 2 3 4 5; do
  echo "== LCGP_DIAG_MODE=$m"
  DIAG_ONLY=1 LCGP_DIAG_MODE=$m timeout 120 python tools/potrf_microbench.py
done
