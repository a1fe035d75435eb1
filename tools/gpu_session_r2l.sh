#!/bin/bash
# Round-2 GPU session L (1 GPU): final validation: full suite, smoke, isolated Cholesky timings, small-problem latencies,
# default bench line, config 5 batch.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2l_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
(echo "# python tools/potrf_microbench.py : lcgp_potrf_batched alone (persistent kernel, factor only)"; timeout 300 python tools/potrf_microbench.py 2>&1 | grep -v "diagonal-block"; echo "# LCGP_POTRF=panels (launch chain, round-1 structure with the new diagonal kernel)"; LCGP_POTRF=panels CASES=1024x1,1024x8,2048x10,4096x4,8064x1,8064x4 timeout 300 python tools/potrf_microbench.py 2>&1 | grep -v "check\|diagonal-block") > gpurun_out/r2l_potrf_microbench.txt; cat gpurun_out/r2l_potrf_microbench.txt
(timeout 200 python tools/cfg5_latency.py cfg5_one; timeout 200 python tools/cfg5_latency.py cfg3_rep; timeout 200 python tools/stage_times.py cfg5_one 64) 2>&1 | grep -v Warn > gpurun_out/r2l_small_problem_latency.txt; cat gpurun_out/r2l_small_problem_latency.txt
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/r2l_bench_1gpu.json 2> gpurun_out/r2l_bench_1gpu.err; echo "bench rc=$?"; python - <<'PY'
import json
j = json.loads([l for l in open('gpurun_out/r2l_bench_1gpu.json') if l.startswith('{')][0])
print('value', j['value'], 'ms/step', j['ms_per_step'], 'e2e', j['e2e']['value'], 'launches', j['gpu_launches'], 'ctor', j['ctor_s'])
print('fit', j.get('fit')); print('predict', j['predict']['wall_ms'])
for k, v in j.get('named_configs', {}).items(): print(k, round(v['ms_per_eval'], 3), 'ms', round(v.get('speedup_vs_cpu_port', 0), 1), 'x')
PY
timeout 900 python bench.py --config cfg5_batch --emulators 64 > gpurun_out/r2l_cfg5_1gpu.json 2> gpurun_out/r2l_cfg5_1gpu.err; echo "cfg5 rc=$?"; cut -c1-330 gpurun_out/r2l_cfg5_1gpu.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2l_bench_reference.json 2>/dev/null; cut -c1-400 gpurun_out/r2l_bench_reference.json
