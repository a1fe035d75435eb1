"""How far do the CUDA path and the CPU oracle drift apart under one shared optimizer? (developer probe)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
from lcgp_b200 import LCGP, synthetic
from oracle import lcgp_oracle as O
from helpers import make_ragged_rep_data
torch.set_num_threads(4)
r = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
cases = {'ragged n=80 (noisy, weakly identified)': (make_ragged_rep_data(seed=9, n_unique=80, p=4, d=2)[:2], dict(q=3, submethod='rep')),
         'notebook case 2': (synthetic.rep1d_skewed()[:2], dict(q=3, submethod='rep')),
         'rep3d n=200': (synthetic.rep3d()[:2], dict(q=3, submethod='rep'))}
for name, ((x, y), mk) in cases.items():
    for maxiter in (10, 30, 1000):
        m = LCGP(y=y, x=x, **mk); o = O.LCGPOracle(y=y, x=x, **mk)
        m.fit(maxiter=maxiter); o.fit(maxiter=maxiter)
        print(f'{name:40s} maxiter={maxiter:4d} evals {m.n_evals:4d}/{o.opt_result.nfev:4d} loss {float(m.loss()):.12f} / {float(o.loss().detach()):.12f} '
              f'lLmb {r(m.lLmb.numpy(), o.lLmb.detach().numpy()):.1e} lLmb0 {r(m.lLmb0.numpy(), o.lLmb0.detach().numpy()):.1e} '
              f'lsig {r(m.lsigma2s.numpy(), o.lsigma2s.detach().numpy()):.1e} lnug {r(m.lnugGPs.numpy(), o.lnugGPs.detach().numpy()):.1e}')
