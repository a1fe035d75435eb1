#!/bin/bash
CASES=128x1,128x32,1024x1,1024x8,2048x10,8064x1,8064x4 timeout 120 python tools/potrf_microbench.py
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/cfg5_latency.py cfg5_one
python tools/cfg5_latency.py cfg3_rep
