#!/bin/bash
# Round-2 GPU session C: full suite + bench with the persistent Cholesky kernel; stream-group sweep at q_loc = 4 and 32.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2c_pytest_gpu.log
run() {  # cfg, env assignments...
  cfg=$1; shift
  echo -n "$cfg $* : "
  env "$@" timeout 300 python bench.py --config $cfg --steps 4 --warmup 3 --no-cpu-baseline --no-fit 2>/dev/null | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        j = json.loads(ln); s = j['stages']
        print('ms/step %.2f  chol %.2f (%.3f) trtri %.2f contract %.2f launches %d' % (j['ms_per_step'], s['cholesky_ms'], s['cholesky_frac_of_dgemm'], s['trtri_ms'], s['contract_kernel_ms'], j['gpu_launches']))
"
}
run cfg4_shard8 X=1
run cfg4_shard8 LCGP_STREAMS=2
run cfg4_shard8 LCGP_STREAMS=4
run cfg4_shard8 LCGP_PLL_HALF=300
run cfg4_shard8 LCGP_POTRF=panels
run cfg4_rep X=1
run cfg4_rep LCGP_STREAMS=1
run cfg4_rep LCGP_STREAMS=2
run cfg4_rep LCGP_POTRF=panels
