#!/bin/bash
# Round-2 GPU session M (1 GPU): structural-zero skipping (padded last block, triangular diagonal blocks) A/B against
# the previous build (lcgp_b200/_lib/liblcgp_b200_base.so), then the full GPU suite.
mkdir -p gpurun_out
BASE=$PWD/lcgp_b200/_lib/liblcgp_b200_base.so
O=gpurun_out/r2m_stage_times.txt; : > $O
for cfg in "cfg4_rep 32" "cfg4_rep 4" "cfg3_rep 10" "cfg5_one 8" "cfg5_one 64"; do
  echo "# new: $cfg" >> $O; timeout 300 python tools/stage_times.py $cfg 2>&1 | grep -v Warn >> $O
  if [ -f $BASE ]; then echo "# base: $cfg" >> $O; LCGP_B200_LIB=$BASE timeout 300 python tools/stage_times.py $cfg 2>&1 | grep -v Warn >> $O; fi
done
cat $O
(timeout 200 python tools/cfg5_latency.py cfg5_one; timeout 200 python tools/cfg5_latency.py cfg3_rep) 2>&1 | grep -v Warn > gpurun_out/r2m_small_problem_latency.txt; cat gpurun_out/r2m_small_problem_latency.txt
(CASES=1024x1,1024x8,2048x10,4096x4,8064x1,8064x4 timeout 300 python tools/potrf_microbench.py 2>&1 | grep -v "diagonal-block") > gpurun_out/r2m_potrf_microbench.txt; cat gpurun_out/r2m_potrf_microbench.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2m_pytest_gpu.log
