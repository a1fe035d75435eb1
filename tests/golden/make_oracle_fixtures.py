"""Regenerates tests/golden/oracle_fixtures.json: small input/output vectors of the CPU oracle.

The reference (TensorFlow/GPflow/TFP) cannot be imported in this image, so these are outputs of the
oracle restatement (oracle/lcgp_oracle.py), whose credibility rests on tests/test_oracle_golden.py
(notebook goldens).  They guard the oracle against accidental edits and give the CUDA parity tests a
committed known answer that does not depend on running the oracle on the GPU box.

    python tests/golden/make_oracle_fixtures.py
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lcgp_b200 import synthetic  # noqa: E402
from oracle.lcgp_oracle import LCGPOracle  # noqa: E402


def case(name, x, y, x0, **mk):
    o = LCGPOracle(y=y, x=x, **mk)
    # 'full' cases: the Cholesky form of lcgp.py:635-666 (neglpost_chol; equals the literal eigh form to ~1e-15 at
    # well-conditioned points, tests/test_oracle_selfcheck.py) -- the eigh form itself loses digits through 1/W
    fn = o.neglpost_chol if mk.get('submethod', 'full') == 'full' else None
    f0, g0 = o.loss_and_grad(fn)
    rng = np.random.default_rng(11)
    q, d = int(o.q), int(o.d)
    lL = o.lLmb.detach().numpy() * rng.uniform(0.6, 1.6, (q, d))
    l0 = rng.uniform(0.5, 3.0, q)
    ls = o.lsigma2s.detach().numpy() + rng.normal(0, 0.3, o.lsigma2s.numel())
    ln = np.exp(rng.uniform(-12, -5, q))
    o.set_constrained(lL, l0, ls, ln)
    f1, g1 = o.loss_and_grad(fn)
    yp, ypv, ycv = o.predict(torch.as_tensor(x0))
    return {'name': name, 'model': mk, 'init': {'loss': f0, 'grad': g0.tolist()},
            'moved': {'lLmb': lL.tolist(), 'lLmb0': l0.tolist(), 'lsigma2s': ls.tolist(), 'lnugGPs': ln.tolist(),
                      'loss': f1, 'grad': g1.tolist(), 'ypred': yp.numpy().tolist(),
                      'ypredvar': ypv.numpy().tolist(), 'yconfvar': ycv.numpy().tolist()}}


def main():
    torch.set_num_threads(1)
    out = []
    x, y, xt, _ = synthetic.rep1d_skewed()
    out.append(case('rep1d_skewed_q3', x, y, xt[::40], q=3, submethod='rep'))
    out.append(case('rep1d_skewed_q2', x, y, xt[::40], q=2, submethod='rep'))
    x, y, x0 = synthetic.rep3d()
    out.append(case('rep3d', x, y, x0[:8], q=3, submethod='rep'))
    x, y, x0, _ = synthetic.latent_mixture(n=150, d=4, p=6, q_true=3, seed=5, rep_choices=None, n0=8)
    out.append(case('full_n150', x, y, x0, q=3, submethod='full'))
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'oracle_fixtures.json'), 'w') as f:
        json.dump(out, f)
    print('wrote', len(out), 'cases')


if __name__ == '__main__':
    main()
