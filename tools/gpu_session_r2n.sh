#!/bin/bash
# A/B of library variants (developer tool): stage times of the config-4 per-GPU share and of batched small matrices,
# isolated Cholesky timings.   VARIANTS="base new" bash tools/gpu_session_r2n.sh
mkdir -p gpurun_out
O=gpurun_out/r2n_variants.txt; : > $O
for v in ${VARIANTS:-base new}; do
  lib=$PWD/lcgp_b200/_lib/liblcgp_b200_$v.so; [ $v = new ] && lib=$PWD/lcgp_b200/_lib/liblcgp_b200.so
  [ -f $lib ] || continue
  echo "## variant $v" >> $O
  for cfg in ${CFGS:-cfg4_rep:4 cfg5_one:64 cfg3_rep:10}; do
    LCGP_B200_LIB=$lib timeout 300 python tools/stage_times.py ${cfg%%:*} ${cfg##*:} 2>&1 | grep -v "Warn\|TFLOP" >> $O
  done
  LCGP_B200_LIB=$lib CASES=${CASES:-1024x8,4096x4,8064x4} timeout 300 python tools/potrf_microbench.py 2>&1 | grep -v "diagonal-block\|check" >> $O
done
cat $O
