"""Shared test utilities: small data sets and an oracle-backed engine.

`OracleEngine` implements the engine interface of lcgp_b200.model (evaluate / predict_latents / aux)
with the CPU oracle's arithmetic.  It exists so that the HOST logic of the sharded path (latent
partition, flat-vector assembly, all-reduce, gather) can be exercised on CPU ranks with gloo; it is
test infrastructure and is never importable from the product package.
"""
import numpy as np
import torch

from oracle.lcgp_oracle import Matern32 as oracle_matern

DT = torch.float64


def make_rep_data(seed=0, n_unique=20, p=4, d=2, reps=3):
    rng = np.random.default_rng(seed)
    xu = rng.uniform(0, 1, (n_unique, d))
    return np.tile(xu, (reps, 1)), rng.standard_normal((p, n_unique * reps)), xu


def make_ragged_rep_data(seed=0, n_unique=60, p=5, d=3):
    rng = np.random.default_rng(seed)
    xu = rng.uniform(0, 1, (n_unique, d))
    r = rng.integers(1, 5, n_unique)
    x = np.repeat(xu, r, axis=0)
    y = np.sin(x @ rng.standard_normal((d, p))).T + 0.1 * rng.standard_normal((p, x.shape[0]))
    return x, y, xu


def make_full_data(seed=0, n=50, p=4, d=2):
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, 1, (n, d))
    y = np.sin(x @ rng.standard_normal((d, p))).T + 0.1 * rng.standard_normal((p, n))
    return x, y


def move_params(model, oracle=None, seed=11):
    """Deterministic off-init parameter point, applied to the model and (optionally) the oracle."""
    rng = np.random.default_rng(seed)
    q, d = int(model.q), int(model.d)
    lL = model.lLmb.numpy() * rng.uniform(0.6, 1.6, (q, d))
    l0 = rng.uniform(0.5, 3.0, q)
    ls = model.lsigma2s.numpy() + rng.normal(0, 0.3, model.lsigma2s.numpy().size)
    ln = np.exp(rng.uniform(-12, -5, q))
    model.lLmb.assign(lL); model.lLmb0.assign(l0); model.lsigma2s.assign(ls); model.lnugGPs.assign(ln)
    if oracle is not None:
        oracle.set_constrained(lL, l0, ls, ln)
    return lL, l0, ls, ln


class OracleEngine:
    """CPU stand-in for CudaEngine: same constructor keywords, same `out` vector layout."""

    def __init__(self, n, d, p, X, sr, YR, w, t, phi_loc, D_loc, scale, sum_log_r, include_host_terms):
        self.n, self.d, self.p = n, d, p
        self.X, self.sr, self.YR, self.w, self.t = X, sr, YR, w, t
        self.phi, self.D = phi_loc, D_loc
        self.q_loc = phi_loc.shape[1]
        self.scale, self.sum_log_r, self.host = scale, sum_log_r, include_host_terms
        self._alpha = self._m = self._L = None

    def _terms(self, lLmb, lLmb0, lnug, lsig_p):
        n, p = self.n, self.p
        s = torch.exp(-0.5 * lsig_p) * self.t
        tot = torch.zeros((), dtype=DT)
        alphas, ms, Ls, logdets, quads = [], [], [], [], []
        eye = torch.eye(n, dtype=DT)
        r = self.sr ** 2
        for k in range(self.q_loc):
            C = oracle_matern(self.X, self.X, lLmb[k], lLmb0[k], lnug[k])
            b = self.YR.T @ (s * self.phi[:, k])
            A = eye + self.D[k] * ((C * self.sr[None, :]) * self.sr[:, None])
            L = torch.linalg.cholesky(A)
            at = torch.cholesky_solve((b / self.sr)[:, None], L)[:, 0]
            alpha = self.sr * at
            m = C @ alpha
            ld = 2.0 * torch.log(torch.diagonal(L)).sum()
            qd = b @ m
            tot = tot + 0.5 * ld - 0.5 * qd
            alphas.append(alpha.detach()); ms.append(m.detach()); Ls.append(L.detach())
            logdets.append(ld.detach()); quads.append(qd.detach())
        if self.host:
            tot = tot + 0.5 * (s ** 2 * self.w).sum() + 0.5 * n * (lsig_p - 2 * torch.log(self.t)).sum() \
                - 0.5 * p * self.sum_log_r
        self._alpha, self._m, self._L = torch.stack(alphas), torch.stack(ms), Ls
        return self.scale * tot, torch.stack(logdets), torch.stack(quads)

    def evaluate(self, lLmb, lLmb0, lnug, lsig_p, with_grad=True, events=None):
        with torch.enable_grad():      # called from inside autograd.Function.forward (grad mode off)
            ins = [t.detach().clone().requires_grad_(True) for t in (lLmb, lLmb0, lnug, lsig_p)]
            val, ld, qd = self._terms(*ins)
        q, d, p = self.q_loc, self.d, self.p
        out = torch.zeros(1 + p + q * d + 4 * q, dtype=DT)
        out[0] = val.detach()
        if with_grad:
            val.backward()
            out[1:1 + p] = ins[3].grad
            o = 1 + p
            out[o:o + q * d] = ins[0].grad.reshape(-1)
            out[o + q * d:o + q * d + q] = ins[1].grad
            out[o + q * d + q:o + q * d + 2 * q] = ins[2].grad
        out[1 + p + q * d + 2 * q:1 + p + q * d + 3 * q] = ld
        out[1 + p + q * d + 3 * q:] = qd
        return out

    def predict_latents(self, lLmb, lLmb0, lnug, x0s, same):
        gh, gv = [], []
        for k in range(self.q_loc):
            c0 = oracle_matern(x0s, self.X, lLmb[k], lLmb0[k], lnug[k])
            gh.append(c0 @ self._alpha[k])
            v = torch.linalg.solve_triangular(self._L[k], (c0 * self.sr[None, :]).T, upper=False)
            gv.append(lLmb0[k] - self.D[k] * (v ** 2).sum(dim=0))
        return torch.stack(gh), torch.stack(gv)

    def aux(self):
        return self._alpha, self._m

    def ainv(self, k):
        return torch.cholesky_inverse(self._L[k])
