"""Where does a config-5 evaluation (n=1024, q=8) spend its time?  (developer tool)
 a) lcgp_plan_run alone in a loop (GPU latency of one graph replay incl. the host sync)
 b) LCGP.loss_and_grad() in a loop (adds the Python host side)
 c) stage events of one launch-by-launch evaluation"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lcgp_b200 import LCGP, synthetic, _cabi
cfg = sys.argv[1] if len(sys.argv) > 1 else 'cfg5_one'
x, y, _, _, mk = synthetic.make_config(cfg)
m = LCGP(y=y, x=x, shard=False, stream_groups=1 if cfg == 'cfg5_one' else 0, **mk)
for _ in range(5): m.loss_and_grad()
eng = m.engine
N = 200
if eng.use_plans:
    plan = eng._plans[1 | eng.group_flags]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(N): eng.lib.lcgp_plan_run(plan, _cabi.stream_ptr())
    t1 = time.perf_counter()
    print(f'{cfg}: plan_run alone {1e3 * (t1 - t0) / N:.3f} ms / eval  (graph: {eng.plan_is_graph()})')
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(N): m.loss_and_grad()
t1 = time.perf_counter()
print(f'{cfg}: loss_and_grad     {1e3 * (t1 - t0) / N:.3f} ms / eval')
lLmb, lLmb0, lsig_p, lnug = m.get_param()
evs = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
for e in evs: e.record()
eng.evaluate_device(lLmb, lLmb0, lnug, lsig_p, True, evs); torch.cuda.synchronize()
eng.evaluate_device(lLmb, lLmb0, lnug, lsig_p, True, evs); torch.cuda.synchronize()
names = ['build', 'cholesky', 'trtri', 'solve', 'contract', 'tail']
print(f'{cfg}: stages (launch by launch, ms): ' + ', '.join(f'{n} {evs[i].elapsed_time(evs[i + 1]):.3f}' for i, n in enumerate(names)))
