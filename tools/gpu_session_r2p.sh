#!/bin/bash
# ncu full captures of the persistent Cholesky kernel and the contraction kernel at the per-GPU share of config 4 (developer tool)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"potrf_pll|gemm_tma" -c 2 -f \
   -o gpurun_out/r2p_shard8 python tools/ncu_eval.py cfg4_shard8 1 > gpurun_out/r2p.log 2>&1; tail -1 gpurun_out/r2p.log
