#!/bin/bash
# A/B of the Cholesky look-ahead (developer tool): isolated potrf timings, then the q_loc=4 shard of config 4.
mkdir -p gpurun_out
for la in 0 1; do
  echo "== LCGP_LOOKAHEAD=$la potrf microbench" 
  LCGP_LOOKAHEAD=$la timeout 300 python tools/potrf_microbench.py
done
for la in 0 1; do for st in 1 2 4; do
  echo "== shard8 LCGP_LOOKAHEAD=$la LCGP_STREAMS=$st"
  LCGP_LOOKAHEAD=$la LCGP_STREAMS=$st timeout 300 python bench.py --config cfg4_shard8 --steps 5 --warmup 3 --no-cpu-baseline --no-fit | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        j = json.loads(ln); s = j['stages']
        print('ms/step %.2f  chol %.2f (%.3f of dgemm) trtri %.2f contract %.2f  obj %.12f' % (j['ms_per_step'], s['cholesky_ms'], s['cholesky_frac_of_dgemm'], s['trtri_ms'], s['contract_kernel_ms'], j['objective']))
"
done; done
