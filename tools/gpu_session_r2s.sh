#!/bin/bash
# Round-2 GPU session S (N GPUs = $1): headline bench at N (with the sharded config-4 fit at N = 8), config 5 at N.
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
EXTRA=""; [ "$N" != 8 ] && EXTRA="--no-fit"
timeout 900 $TR --nproc-per-node $N --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 $EXTRA > gpurun_out/r2s_bench_${N}gpu.json 2> gpurun_out/r2s_bench_${N}gpu.err; echo "bench$N rc=$?"
timeout 600 $TR --nproc-per-node $N --master-port 29541 bench.py --gpus $N --config cfg5_batch --emulators 64 > gpurun_out/r2s_cfg5_${N}gpu.json 2> gpurun_out/r2s_cfg5_${N}gpu.err; echo "cfg5 $N rc=$?"
if [ "$N" = 2 ]; then timeout 600 python -m pytest tests/test_nccl_sharded.py -m gpu -x -q -s > gpurun_out/r2s_pytest_nccl_w2.log 2>&1; echo "nccl rc=$?"; grep -E "NCCL_SHARDED|world=|passed|failed" gpurun_out/r2s_pytest_nccl_w2.log | tail -6; fi
python - <<PY
import json
for f in ['gpurun_out/r2s_bench_${N}gpu.json']:
    ls = [l for l in open(f) if l.startswith('{')]
    if not ls: print(f, 'NO LINE'); continue
    j = json.loads(ls[0]); s = j['stages']
    print(f, 'N', j['n_gpus'], 'value %.3f ms/step %.2f e2e %.3f' % (j['value'], j['ms_per_step'], j['e2e']['value']), 'launches', j['gpu_launches'],
          'chol %.2f (%.3f, fused %s) trtri %.2f contract %.2f' % (s['cholesky_ms'], s['cholesky_frac_of_dgemm'], s['cholesky_trtri_fused'], s['trtri_ms'], s['contract_kernel_ms']),
          'roofline %.3f %s' % (j['roofline']['frac'], j['roofline']['kernel'][:20]), 'predict', round(j['predict']['wall_ms'], 1), 'ctor', round(j['ctor_s'], 2), 'fit', j.get('fit'))
for f in ['gpurun_out/r2s_cfg5_${N}gpu.json']:
    ls = [l for l in open(f) if l.startswith('{')]
    if not ls: print(f, 'NO LINE'); continue
    j = json.loads(ls[0]); print(f, 'N', j['n_gpus'], 'fits/s %.2f evals/s %.0f wall %.2f conv %d' % (j['value'], j['evals_per_s'], j['wall_s'], j['converged']))
PY
