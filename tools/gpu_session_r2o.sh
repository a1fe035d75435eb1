#!/bin/bash
# ncu full captures (source-level) of one kernel for several library variants (developer tool)
#   VARIANTS="C new" KREGEX=gemm_tma bash tools/gpu_session_r2o.sh
mkdir -p gpurun_out
for v in ${VARIANTS:-base new}; do
  lib=$PWD/lcgp_b200/_lib/liblcgp_b200_$v.so; [ $v = new ] && lib=$PWD/lcgp_b200/_lib/liblcgp_b200.so
  LCGP_B200_LIB=$lib timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
     -k regex:${KREGEX:-gemm_tma} -c 1 -f -o gpurun_out/r2o_${KREGEX:-gemm_tma}_$v python tools/ncu_eval.py ${CFG:-cfg4_shard8} 1 > gpurun_out/r2o_$v.log 2>&1
  tail -1 gpurun_out/r2o_$v.log
done
