#!/bin/bash
# Round-2 GPU session E: full suite (device-side predict maps, batched engine with active-set compaction), cfg5 trace.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2e_pytest_gpu.log
LCGP_BATCH_TRACE=1 timeout 900 python bench.py --config cfg5_batch --emulators 64 --engine lockstep --no-cpu-baseline 2>&1 | grep -v Warning | cut -c1-500
LCGP_BATCH_TRACE=1 timeout 900 python bench.py --config cfg5_batch --emulators 8 --engine lockstep --no-cpu-baseline 2>&1 | grep -v Warning | cut -c1-400
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-fit 2>/dev/null | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        j = json.loads(ln); print('cfg4 ms/step', j['ms_per_step'], 'predict', j['predict'])"
