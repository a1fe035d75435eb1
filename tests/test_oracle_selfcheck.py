"""Self-consistency of the CPU oracle (SURVEY 8c 'self-consistency oracles') and of the analytic
formulas the CUDA path implements, all on the CPU."""
import json
import os

import numpy as np
import torch

from lcgp_b200 import synthetic
from oracle.lcgp_oracle import LCGPOracle, Matern32, SoftClip
from helpers import make_full_data, make_ragged_rep_data, move_params

FIX = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'oracle_fixtures.json')))


def _rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def test_fixture_regression():
    """The committed oracle fixtures are reproduced (guards the oracle against accidental edits)."""
    x, y, xt, _ = synthetic.rep1d_skewed()
    case = next(c for c in FIX if c['name'] == 'rep1d_skewed_q3')
    o = LCGPOracle(y=y, x=x, **case['model'])
    f, g = o.loss_and_grad()
    assert abs(f - case['init']['loss']) <= 1e-12 * abs(f)
    assert _rel(g, case['init']['grad']) < 1e-9
    mv = case['moved']
    o.set_constrained(mv['lLmb'], mv['lLmb0'], mv['lsigma2s'], mv['lnugGPs'])
    f, g = o.loss_and_grad()
    assert abs(f - mv['loss']) <= 1e-12 * abs(f)
    yp, _, _ = o.predict(torch.as_tensor(xt[::40]))
    assert _rel(yp.numpy(), mv['ypred']) < 1e-10


def test_softclip_roundtrip_and_bounds():
    for lo, hi in [(1e-6, 1e4), (1e-4, 1e4), (np.exp(-16.0), np.exp(-2.0))]:
        b = SoftClip(lo, hi)
        y = torch.as_tensor(np.geomspace(lo * 1.5, hi * 0.5, 9))
        assert _rel(b.forward(b.inverse(y)).numpy(), y.numpy()) < 1e-9
        u = torch.linspace(-50, 50, 11, dtype=torch.float64)
        v = b.forward(u)
        # the wide intervals pass through intermediates of size 1e4: ~1e-12 absolute rounding noise
        assert float(v.min()) >= lo - 1e-11 and float(v.max()) <= hi + 1e-11


def test_matern_properties():
    rng = np.random.default_rng(0)
    x = torch.as_tensor(rng.uniform(0, 1, (30, 3)))
    ell = torch.as_tensor([0.3, 0.7, 1.1]); s0 = torch.tensor(2.0); nug = torch.tensor(1e-3)
    C = Matern32(x, x, ell, s0, nug)
    assert torch.allclose(C, C.T)
    assert torch.allclose(torch.diagonal(C), s0 * torch.ones(30, dtype=torch.float64))    # (1-nu) + nu = 1
    assert float(torch.linalg.eigvalsh(C).min()) > 0
    x2 = x[:7]
    C2 = Matern32(x2, x, ell, s0, nug)                      # unequal shapes: no nugget (covmat.py:46-47)
    nu = nug / (1 + nug)
    assert torch.allclose(C2[:, 7:], C[:7, 7:])
    assert torch.allclose(torch.diagonal(C2[:, :7]), (1 - nu) * s0 * torch.ones(7, dtype=torch.float64))


def test_full_loss_eigh_form_equals_cholesky_form():
    x, y = make_full_data(seed=3, n=60, p=5, d=3)
    o = LCGPOracle(y=y, x=x, q=3, submethod='full')
    move_params(o_model := _Shim(o), None)
    a = float(o.neglpost().detach()); b = float(o.neglpost_chol().detach())
    assert abs(a - b) <= 1e-12 * abs(a)
    _, ga = o.loss_and_grad(o.neglpost); _, gb = o.loss_and_grad(o.neglpost_chol)
    assert _rel(ga, gb) < 1e-8          # eigh backward is the less accurate of the two


class _Shim:
    """Lets helpers.move_params drive an oracle (it expects .numpy()/.assign on parameters)."""

    def __init__(self, o):
        self.o, self.q, self.d = o, o.q, o.d
        mk = lambda get, idx: type('P', (), {'numpy': staticmethod(lambda: get().detach().numpy()),
                                             'assign': staticmethod(lambda v: self._set(idx, v))})()
        self.lLmb = mk(lambda: o.lLmb, 0); self.lLmb0 = mk(lambda: o.lLmb0, 1)
        self.lsigma2s = mk(lambda: o.lsigma2s, 2); self.lnugGPs = mk(lambda: o.lnugGPs, 3)

    def _set(self, idx, v):
        cur = [self.o.lLmb, self.o.lLmb0, self.o.lsigma2s, self.o.lnugGPs]
        cur = [c.detach().numpy().copy() for c in cur]
        cur[idx] = np.asarray(v)
        self.o.set_constrained(*cur)


def test_analytic_gradient_formulas_vs_autograd():
    """The single-factor identities and analytic gradient of SURVEY A.4/A.5 (what csrc/solve_grad.cu
    computes), evaluated in numpy, against autograd through the literal restatement."""
    x, y, _ = make_ragged_rep_data(seed=0, n_unique=60, p=6, d=3)
    es = [2, 1, 3]
    o = LCGPOracle(y=y, x=x, q=4, submethod='rep', diag_error_structure=es)
    move_params(_Shim(o), None)
    val, gl, gs0, gsig, gnug = o.grad_constrained()
    X = o.x_unique_s.numpy(); rr = o.r.numpy().astype(float); sr = np.sqrt(rr); n = o.n; p = o.p; q = o.q; d = o.d
    ell = o.lLmb.detach().numpy(); s0 = o.lLmb0.detach().numpy(); lnug = o.lnugGPs.detach().numpy()
    lsig_p = o.get_param()[2].detach().numpy(); t = o.ybar_std[:, 0].numpy(); Y = o.ybar_s.numpy()
    phi = o.phi.numpy(); D = o.diag_D.numpy()
    s = np.exp(-0.5 * lsig_p) * t; YR = Y * rr[None, :]; w = (YR * Y).sum(1)
    F = 0.5 * (s ** 2 * w).sum() + n / 2 * (lsig_p - 2 * np.log(t)).sum() - p / 2 * np.log(rr).sum()
    g_l = np.zeros((q, d)); g_s = np.zeros(q); g_n = np.zeros(q); g_sig = -0.5 * s ** 2 * w + n / 2
    for k in range(q):
        Sj = [np.abs(X[:, j][:, None] - X[:, j][None, :]) / ell[k, j] for j in range(d)]
        C0 = np.prod([1 + S for S in Sj], axis=0) * np.exp(-sum(Sj)); nu = lnug[k] / (1 + lnug[k])
        C = s0[k] * ((1 - nu) * C0 + nu * np.eye(n))
        b = YR.T @ (s * phi[:, k])
        A = np.eye(n) + D[k] * sr[:, None] * sr[None, :] * C
        L = np.linalg.cholesky(A); W = np.linalg.inv(L); Ainv = W.T @ W
        alpha = sr * (Ainv @ (b / sr)); mk = (b - alpha) / (D[k] * rr)
        assert _rel(mk, C @ alpha) < 1e-10                       # m = C alpha = (b - alpha)/(d r)
        F += np.log(np.diag(L)).sum() - 0.5 * b @ mk
        G = D[k] * sr[:, None] * sr[None, :] * Ainv - np.outer(alpha, alpha)
        g_s[k] = 0.5 * (G * C).sum() / s0[k]
        g_n[k] = 0.5 * (G * (s0[k] * (np.eye(n) - C0))).sum() / (1 + lnug[k]) ** 2
        for j in range(d):
            g_l[k, j] = 0.5 * (G * (s0[k] * (1 - nu) * C0 * Sj[j] ** 2 / ((1 + Sj[j]) * ell[k, j]))).sum()
        g_sig += 0.5 * s * phi[:, k] * (YR @ mk)
    F /= n
    g_sig_grp = np.add.reduceat(g_sig / n, np.cumsum([0] + es[:-1]))
    assert abs(F - val) <= 1e-13 * abs(val)
    assert _rel(g_l / n, gl.numpy()) < 1e-11
    assert _rel(g_s / n, gs0.numpy()) < 1e-11
    assert _rel(g_n / n, gnug.numpy()) < 1e-11
    assert _rel(g_sig_grp, gsig.numpy()) < 1e-11


def test_stable_and_literal_rep_aux_agree():
    """Tk through explicit inv(C) (lcgp.py:783-788) vs d (sqrt r sqrt r^T) o A^-1: same operator; the
    literal form is the inaccurate one (cond(C) ~ 1/nugget), so only 1e-6 is asked of it."""
    x, y, _ = make_ragged_rep_data(seed=2, n_unique=40, p=4, d=2)
    o = LCGPOracle(y=y, x=x, q=3, submethod='rep')
    o._compute_aux_predictive_quantities_rep(stable=True); T1 = o.Tks.clone(); c1 = o.CinvMs.clone()
    o._compute_aux_predictive_quantities_rep(stable=False); T2 = o.Tks.clone()
    assert _rel(T2.numpy(), T1.numpy()) < 1e-6
    # CinvM_k == sqrt(r) o A^-1 (b / sqrt r)  is what the CUDA path stores as alpha
    assert torch.isfinite(c1).all()


def test_finite_difference_gradient():
    x, y, _ = make_ragged_rep_data(seed=4, n_unique=30, p=3, d=2)
    o = LCGPOracle(y=y, x=x, q=2, submethod='rep')
    f0, g = o.loss_and_grad()
    v0 = o._flat_get()
    rng = np.random.default_rng(0)
    for _ in range(3):
        dirn = rng.standard_normal(v0.size); dirn /= np.linalg.norm(dirn)
        h = 1e-6
        o._flat_set(v0 + h * dirn); fp = float(o.loss().detach())
        o._flat_set(v0 - h * dirn); fm = float(o.loss().detach())
        assert abs((fp - fm) / (2 * h) - g @ dirn) <= 1e-6 * max(1.0, abs(g @ dirn))
    o._flat_set(v0)
