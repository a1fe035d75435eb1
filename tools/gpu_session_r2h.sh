#!/bin/bash
# Round-2 GPU session H: ncu launch lists and full captures of the new / dominant kernels (one evaluation each).
mkdir -p gpurun_out
set -x
python tools/ncu_eval.py cfg4_shard8 2 > gpurun_out/r2h_plain_shard8.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2h_launches_cfg4_shard8.csv python tools/ncu_eval.py cfg4_shard8 2 > gpurun_out/r2h_ncu1.log 2>&1
python tools/ncu_eval.py cfg3_rep 2 > gpurun_out/r2h_plain_cfg3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2h_launches_cfg3_rep.csv python tools/ncu_eval.py cfg3_rep 2 > gpurun_out/r2h_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"potrf_pll_kernel|ContractJob|build_A_kernel" -c 3 -o gpurun_out/r2h_prof_shard8 python tools/ncu_eval.py cfg4_shard8 2 > gpurun_out/r2h_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"potrf_pll_kernel|ContractJob" -c 2 -o gpurun_out/r2h_prof_cfg3 python tools/ncu_eval.py cfg3_rep 2 > gpurun_out/r2h_ncu4.log 2>&1
python tools/ncu_eval.py cfg4_rep 1 > gpurun_out/r2h_plain_cfg4.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2h_launches_cfg4_rep.csv python tools/ncu_eval.py cfg4_rep 1 > gpurun_out/r2h_ncu5.log 2>&1
ls -la gpurun_out/r2h_*
tail -3 gpurun_out/r2h_ncu3.log gpurun_out/r2h_ncu4.log
