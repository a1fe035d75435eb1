#!/bin/bash
for h in 0 74; do
  echo "== LCGP_HALF_TILES=$h"
  CASES=256x1,1024x1,1024x8,2048x10,8064x1,8064x4 LCGP_HALF_TILES=$h timeout 120 python tools/potrf_microbench.py
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/cfg5_latency.py cfg5_one
python tools/cfg5_latency.py cfg3_rep
