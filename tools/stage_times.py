"""Stage times (CUDA events recorded inside lcgp_nll_grad) for a configuration with the latent count overridden:
    python tools/stage_times.py cfg5_one 64        (developer tool)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lcgp_b200 import LCGP, synthetic
cfg = sys.argv[1]; q = int(sys.argv[2])
x, y, _, _, mk = synthetic.make_config(cfg)
mk['q'] = q
m = LCGP(y=y, x=x, shard=False, **mk)
eng = m.engine
lLmb, lLmb0, lsig_p, lnug = m.get_param()
evs = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
for e in evs: e.record()
for _ in range(3):
    eng.evaluate_device(lLmb, lLmb0, lnug, lsig_p, True, evs); torch.cuda.synchronize()
names = ['build', 'cholesky', 'trtri', 'solve', 'contract', 'tail']
t = [evs[i].elapsed_time(evs[i + 1]) for i in range(6)]
n = int(m.n); fl = q * n ** 3 / 3
print(f'{cfg} q={q} n={n}: ' + ', '.join(f'{nm} {v:.3f}' for nm, v in zip(names, t)) + f' | total {sum(t):.3f} ms')
print('   TFLOP/s: cholesky %.1f trtri %.1f contract %.1f' % (fl / t[1] / 1e9, fl / t[2] / 1e9, fl / t[4] / 1e9))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): eng.evaluate_device(lLmb, lLmb0, lnug, lsig_p, True)
e1.record(); torch.cuda.synchronize()
print('   without stage events (stream groups free to overlap): %.3f ms per evaluation' % (e0.elapsed_time(e1) / 5))
