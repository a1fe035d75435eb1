#!/bin/bash
# Round-2 GPU session G: tuning of the fused persistent Cholesky + inverse (tile heights, stream groups) per configuration.
lat() { echo -n "$* : "; env "$@" timeout 300 python tools/cfg5_latency.py $CFG 2>&1 | grep -E "plan_run|stages" | sed 's/.*plan_run alone/plan/; s/.*stages (launch by launch, ms)://' | tr '\n' ' '; echo; }
CFG=cfg3_rep
lat X=1
lat LCGP_PLL_HALF=0
lat LCGP_PLL_QUARTER=200
CFG=cfg5_one
lat X=1
lat LCGP_PLL_QUARTER=0
lat LCGP_PLL_QUARTER=0 LCGP_PLL_HALF=0
run() {  # cfg, env assignments...
  cfg=$1; shift
  echo -n "$cfg $* : "
  env "$@" timeout 300 python bench.py --config $cfg --steps 4 --warmup 3 --no-cpu-baseline --no-fit 2>/dev/null | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        j = json.loads(ln); s = j['stages']
        print('ms/step %.2f  chol %.2f trtri %.2f contract %.2f launches %d' % (j['ms_per_step'], s['cholesky_ms'], s['trtri_ms'], s['contract_kernel_ms'], j['gpu_launches']))
"
}
run cfg4_shard8 X=1
run cfg4_shard8 LCGP_FUSE_TRTRI=0
run cfg4_shard8 LCGP_PLL_HALF=300
run cfg4_shard8 LCGP_STREAMS=2
run cfg4_rep X=1
run cfg4_rep LCGP_POTRF=pll
run cfg4_rep LCGP_POTRF=pll LCGP_STREAMS=1
