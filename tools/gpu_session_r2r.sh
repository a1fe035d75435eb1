#!/bin/bash
# Round-2 GPU session R (1 GPU): measurements after the structural-zero skipping and the late-fill producer:
# default bench line, isolated Cholesky timings, small-problem latencies, config 5 batch, ncu launch lists.
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/r2r_bench_1gpu.json 2> gpurun_out/r2r_bench_1gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
j = json.loads([l for l in open('gpurun_out/r2r_bench_1gpu.json') if l.startswith('{')][0])
print('value', j['value'], 'ms/step', j['ms_per_step'], 'e2e', j['e2e']['value'], 'launches', j['gpu_launches'], 'ctor', j['ctor_s'], 'roofline', j['roofline']['frac'])
print('stages', {k: round(v, 3) for k, v in j['stages'].items() if isinstance(v, float)})
print('fit', j.get('fit')); print('predict', j['predict']['wall_ms'])
for k, v in j.get('named_configs', {}).items(): print(k, round(v['ms_per_eval'], 3), 'ms', round(v.get('speedup_vs_cpu_port', 0), 1), 'x')
PY
(echo "# python tools/potrf_microbench.py : lcgp_potrf_batched alone (persistent kernel, factor only)"; timeout 300 python tools/potrf_microbench.py 2>&1 | grep -v "diagonal-block") > gpurun_out/r2r_potrf_microbench.txt; cat gpurun_out/r2r_potrf_microbench.txt
(timeout 200 python tools/cfg5_latency.py cfg5_one; timeout 200 python tools/cfg5_latency.py cfg3_rep; timeout 200 python tools/stage_times.py cfg5_one 64; timeout 300 python tools/stage_times.py cfg4_rep 4) 2>&1 | grep -v Warn > gpurun_out/r2r_small_problem_latency.txt; cat gpurun_out/r2r_small_problem_latency.txt
timeout 900 python bench.py --config cfg5_batch --emulators 64 > gpurun_out/r2r_cfg5_1gpu.json 2> gpurun_out/r2r_cfg5_1gpu.err; echo "cfg5 rc=$?"; cut -c1-330 gpurun_out/r2r_cfg5_1gpu.json
for cfg in cfg4_rep cfg4_shard8 cfg3_rep; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2r_launches_$cfg.csv python tools/ncu_eval.py $cfg 1 > /dev/null 2>&1
  python tools/ncu_summary.py gpurun_out/r2r_launches_$cfg.csv > gpurun_out/r2r_launches_$cfg.txt; head -8 gpurun_out/r2r_launches_$cfg.txt
done
