"""Smallest end-to-end case for compute-sanitizer (developer tool): objective + gradient + prediction, n = 300 (3 blocks),
q = 2, persistent Cholesky + fused inverse, then the launch chain, then a batched (2-emulator) evaluation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
from lcgp_b200 import LCGP
from lcgp_b200.batched import BatchedEngine
from helpers import make_ragged_rep_data
os.environ.setdefault('LCGP_GRAPHS', '0')
ms = []
for e in range(2):
    x, y, _ = make_ragged_rep_data(seed=60 + e, n_unique=300, p=5, d=3)
    ms.append(LCGP(y=y, x=x, q=2, submethod='rep'))
f, g = ms[0].loss_and_grad()
out = ms[0].predict(np.random.default_rng(0).uniform(0, 1, (50, 3)))
pars = [m.get_param() for m in ms]
st = lambda k: torch.stack([p[k] for p in pars])
o = BatchedEngine(ms).evaluate(st(0), st(1), st(3), st(2), True)
torch.cuda.synchronize()
print('sanitize case ok', f, float(o[0, 0]), float(o[1, 0]), float(out[0].sum()))
