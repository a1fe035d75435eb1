#!/bin/bash
# Round-end style verification: GPU parity suite, smoke, default bench, reference arm, config-5 batch (developer tool).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; tail -3 gpurun_out/pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_1gpu_final.json 2> gpurun_out/bench_1gpu_final.err; echo "bench rc=$?"; python - <<'PY'
import json
j = json.loads([l for l in open('gpurun_out/bench_1gpu_final.json') if l.startswith('{')][0])
print('value', j['value'], 'ms/step', j['ms_per_step'], 'e2e', j['e2e']['value'], 'launches', j['gpu_launches'], 'roofline', j['roofline']['frac'], 'traffic', j['roofline']['traffic_source'])
print('stages', {k: round(v, 3) for k, v in j['stages'].items() if isinstance(v, float)})
print('clocks', j['clocks']); print('cpu_baseline', j['cpu_baseline'])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-300
