#!/bin/bash
# 128 x 64 contraction tiles (two CTAs per SM) A/B: LCGP_CONTRACT_N64=0 / 1 (developer tool)
mkdir -p gpurun_out
O=gpurun_out/r2q_contract_n64.txt; : > $O
for n64 in 0 1; do
  echo "## LCGP_CONTRACT_N64=$n64" >> $O
  for cfg in ${CFGS:-cfg4_rep:4 cfg4_rep:32 cfg5_one:64 cfg3_rep:10 cfg5_one:8}; do
    LCGP_CONTRACT_N64=$n64 timeout 300 python tools/stage_times.py ${cfg%%:*} ${cfg##*:} 2>&1 | grep -v "Warn\|TFLOP" >> $O
  done
done
cat $O
