#!/bin/bash
# Round-2 GPU session J (8 GPUs): headline bench at N = 8 (with the sharded config-4 fit), N = 2 / 4 quick lines,
# config 5 (64 emulators to convergence) at N = 1 / 2 / 4 / 8, NCCL parity test at world 4.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2j_bench_8gpu.json 2> gpurun_out/r2j_bench_8gpu.err; echo "bench8 rc=$?"
for n in 2 4; do timeout 600 $TR --nproc-per-node $n --master-port 2953$n bench.py --gpus $n --steps 10 --warmup 3 --no-fit > gpurun_out/r2j_bench_${n}gpu.json 2> gpurun_out/r2j_bench_${n}gpu.err; echo "bench$n rc=$?"; done
for n in 8 4 2; do timeout 600 $TR --nproc-per-node $n --master-port 2954$n bench.py --gpus $n --config cfg5_batch --emulators 64 > gpurun_out/r2j_cfg5_${n}gpu.json 2> gpurun_out/r2j_cfg5_${n}gpu.err; echo "cfg5 $n rc=$?"; done
timeout 600 python -m pytest tests/test_nccl_sharded.py -m gpu -x -q -s > gpurun_out/r2j_pytest_nccl_w4.log 2>&1; echo "nccl rc=$?"; grep -E "NCCL_SHARDED|world=|passed|failed" gpurun_out/r2j_pytest_nccl_w4.log | tail -6
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2j_bench_*gpu.json')):
    ls = [l for l in open(f) if l.startswith('{')]
    if not ls: print(f, 'NO LINE'); continue
    j = json.loads(ls[0]); s = j['stages']
    print(f, 'N', j['n_gpus'], 'value %.3f ms/step %.2f e2e %.3f' % (j['value'], j['ms_per_step'], j['e2e']['value']), 'launches', j['gpu_launches'],
          'chol %.2f (%.3f, fused %s) trtri %.2f contract %.2f' % (s['cholesky_ms'], s['cholesky_frac_of_dgemm'], s['cholesky_trtri_fused'], s['trtri_ms'], s['contract_kernel_ms']),
          'roofline %.3f %s' % (j['roofline']['frac'], j['roofline']['kernel'][:20]), 'predict', round(j['predict']['wall_ms'], 1), 'ctor', round(j['ctor_s'], 2), 'fit', j.get('fit'))
for f in sorted(glob.glob('gpurun_out/r2j_cfg5_*gpu.json')):
    ls = [l for l in open(f) if l.startswith('{')]
    if not ls: print(f, 'NO LINE'); continue
    j = json.loads(ls[0]); print(f, 'N', j['n_gpus'], 'fits/s %.2f evals/s %.0f wall %.2f conv %d' % (j['value'], j['evals_per_s'], j['wall_s'], j['converged']))
PY
