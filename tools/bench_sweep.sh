#!/bin/bash
# Developer tool: whole-evaluation timing of config 4 (q_loc = 32 and the q_loc = 4 shard) over Cholesky tunables.
run() {  # cfg, env assignments...
  cfg=$1; shift
  echo -n "$cfg $* : "
  env "$@" timeout 300 python bench.py --config $cfg --steps 4 --warmup 3 --no-cpu-baseline --no-fit 2>/dev/null | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        j = json.loads(ln); s = j['stages']
        print('ms/step %.2f  chol %.2f (%.3f) trtri %.2f contract %.2f' % (j['ms_per_step'], s['cholesky_ms'], s['cholesky_frac_of_dgemm'], s['trtri_ms'], s['contract_kernel_ms']))
"
}
run cfg4_rep LCGP_PANEL_W=8 LCGP_TAIL_N=16 LCGP_TAIL_W=4
run cfg4_rep LCGP_PANEL_W=8 LCGP_TAIL_N=24 LCGP_TAIL_W=4
run cfg4_rep LCGP_PANEL_W=8 LCGP_TAIL_N=24 LCGP_TAIL_W=2
run cfg4_rep LCGP_PANEL_W=16 LCGP_TAIL_N=24 LCGP_TAIL_W=4
run cfg4_rep LCGP_PANEL_W=8 LCGP_TAIL_N=24 LCGP_TAIL_W=4 LCGP_STREAMS=2
run cfg4_shard8 LCGP_PANEL_W=8 LCGP_TAIL_N=24 LCGP_TAIL_W=4
run cfg4_shard8 LCGP_PANEL_W=8 LCGP_TAIL_N=24 LCGP_TAIL_W=2
run cfg4_shard8 LCGP_PANEL_W=4 LCGP_TAIL_N=24 LCGP_TAIL_W=4
run cfg4_shard8 LCGP_PANEL_W=4 LCGP_TAIL_N=24 LCGP_TAIL_W=2
run cfg4_shard8 LCGP_PANEL_W=6 LCGP_TAIL_N=24 LCGP_TAIL_W=3
run cfg4_shard8 LCGP_PANEL_W=4 LCGP_TAIL_N=24 LCGP_TAIL_W=2 LCGP_STREAMS=2
run cfg4_shard8 LCGP_PANEL_W=8 LCGP_TAIL_N=24 LCGP_TAIL_W=2 LCGP_STREAMS=2
