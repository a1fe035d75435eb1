#!/bin/bash
# Round-2 GPU session I (1 GPU): full suite, smoke, default bench line (with named configs + CPU baselines), cfg5 batch.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2i_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/r2i_bench_1gpu.json 2> gpurun_out/r2i_bench_1gpu.err; echo "bench rc=$?"; python - <<'PY'
import json
j = json.loads([l for l in open('gpurun_out/r2i_bench_1gpu.json') if l.startswith('{')][0])
print('value', j['value'], 'ms/step', j['ms_per_step'], 'e2e', j['e2e']['value'], 'launches', j['gpu_launches'])
print('roofline', {k: j['roofline'][k] for k in ('kernel', 'achieved', 'peak', 'frac', 'kernel_ms', 'traffic_source')})
print('stages', {k: round(v, 3) if isinstance(v, float) else v for k, v in j['stages'].items()})
print('cpu', j['cpu_baseline']); print('predict', j['predict']); print('fit', j.get('fit')); print('ctor_s', j['ctor_s'])
for k, v in j.get('named_configs', {}).items(): print(k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items()})
PY
timeout 900 python bench.py --config cfg5_batch --emulators 64 > gpurun_out/r2i_cfg5_1gpu.json 2> gpurun_out/r2i_cfg5_1gpu.err; echo "cfg5 rc=$?"; cut -c1-700 gpurun_out/r2i_cfg5_1gpu.json
