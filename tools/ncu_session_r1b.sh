#!/bin/bash
# ncu evidence for the look-ahead build (one gpurun call): launch list of one evaluation of the default bench
# command, full captures of the auxiliary kernels and of the diagonal-block kernel.
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-fit"
$BENCH > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 3381 -c 844 --csv --log-file gpurun_out/launches_r1b.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
python tools/aux_kernels_bench.py > gpurun_out/aux_kernels.json 2> gpurun_out/aux_kernels.err &&
ncu --set full --clock-control none --import-source on -k regex:fullcov -c 1 -o gpurun_out/prof_fullcov python tools/aux_kernels_bench.py > gpurun_out/ncu_fullcov.log 2>&1
DIAG_ONLY=1 python tools/potrf_microbench.py > gpurun_out/diag_plain.log 2>&1 &&
DIAG_ONLY=1 ncu --set full --clock-control none --import-source on -k regex:potrf_diag -s 5 -c 1 -o gpurun_out/prof_diag python tools/potrf_microbench.py > gpurun_out/ncu_diag.log 2>&1
cat gpurun_out/aux_kernels.json; tail -2 gpurun_out/ncu_launches.log; ls -la gpurun_out/*.ncu-rep
