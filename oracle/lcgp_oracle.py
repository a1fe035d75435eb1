"""
CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  NOT A PRODUCT PATH.

A torch-float64 CPU restatement of the LCGP reference's emulator-fitting path
(the reference itself is TensorFlow/GPflow/TFP and cannot be imported in this
image: tensorflow, tensorflow_probability and gpflow are not installed and there
is no network).  Only `tests/`, `__graft_entry__.smoke()` and the CPU-baseline /
`--impl reference` legs of `bench.py` may import this module.  Nothing under
`lcgp_b200/` imports it; the product fails loudly without its CUDA library.

Parity status: PINNED by the reference's only numerical artefacts, the stored
outputs of `illustration-examples/lcgp-rep-1d-illustration.ipynb` (case 2,
seed 123): `diag_D`, Var(g) (8 printed digits, exact), fitted length-scales and
noise log-variances (~2e-4 relative, optimizer-limited) and the five predictive
metrics (all printed digits).  See `tests/test_oracle_golden.py` and
`tests/golden/`.  The reference's own test-suite holds no numerical
known-answer test for loss / gradient / fit.

Every function cites the reference lines it follows (paths relative to
/root/reference).  Gradients come from torch autograd, mirroring the reference's
TF autodiff (`gpflow.optimizers.Scipy` -> tf.GradientTape, lcgp.py:537-540).

Third-party arithmetic that is not in the reference tree and is restated from
its published definition (pinned versions from pyproject.toml:16-24):
  * tensorflow-probability>=0.25.0 `bijectors.SoftClip` (two-sided, hinge_softness=1)
  * tensorflow-probability>=0.25.0 `stats.percentile(x, 50.0)` with its default
    interpolation='nearest'
  * gpflow>=2.5.0 `optimizers.Scipy().minimize` -> scipy L-BFGS-B with defaults
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch

DT = torch.float64


# ---------------------------------------------------------------------------
# Third-party pieces restated from their published definitions
# ---------------------------------------------------------------------------
def _softplus_inverse(a: torch.Tensor) -> torch.Tensor:
    # log(expm1(a)) written to stay finite for large a
    return a + torch.log(-torch.expm1(-a))


class SoftClip:
    """TFP `bijectors.SoftClip(low, high)` with hinge_softness = 1.

    forward(u) = high - softplus(high - low - softplus(u - low)) * (high-low)/softplus(high-low)
    Call sites in the reference: lcgp.py:184-187, 193-196, 206-209.
    """

    def __init__(self, low: float, high: float):
        self.low = torch.tensor(float(low), dtype=DT)
        self.high = torch.tensor(float(high), dtype=DT)
        width = self.high - self.low
        self.scale = width / torch.nn.functional.softplus(width)

    def forward(self, u: torch.Tensor) -> torch.Tensor:
        sp = torch.nn.functional.softplus
        width = self.high - self.low
        return self.high - sp(width - sp(u - self.low)) * self.scale

    def inverse(self, y: torch.Tensor) -> torch.Tensor:
        width = self.high - self.low
        inner = _softplus_inverse((self.high - y) / self.scale)
        return self.low + _softplus_inverse(width - inner)


def percentile50_nearest(Y: torch.Tensor) -> torch.Tensor:
    """TFP `stats.percentile(Y, 50.0, axis=1, keepdims=True)` with the default
    interpolation='nearest': ascending sort, index round((m-1)*0.5) (half-to-even).
    Call sites: lcgp.py:317-318, 388-389."""
    m = Y.shape[1]
    idx = int(np.round((m - 1) * 0.5))  # numpy rounds half to even, as TF does
    srt, _ = torch.sort(Y, dim=1)
    return srt[:, idx:idx + 1]


# ---------------------------------------------------------------------------
# covmat.py:5-55
# ---------------------------------------------------------------------------
def Matern32(x1, x2, llmb, llmb0, lnug, diag_only: bool = False):
    """Separable (product) Matern-3/2 with nugget mixing -- covmat.py:5-55."""
    assert x1.ndim == 2, 'input x1 should be 2-dimensional, (n_param, dim_param)'   # :18
    assert x2.ndim == 2, 'input x2 should be 2-dimensional, (n_param, dim_param)'   # :19
    assert x1.shape[1] == x2.shape[1], 'the dim_param of input x1 and x2 should be the same.'  # :20
    d = x1.shape[1]

    if diag_only:                                                                    # :23-29
        assert bool(torch.all(torch.abs(x1 - x2) <= (1e-6 + 1e-6 * torch.abs(x2)))), \
            'diag_only should only be called when x1 and x2 are identical.'
        return llmb0 * torch.ones(x1.shape[0], dtype=DT)

    V = torch.zeros((x1.shape[0], x2.shape[0]), dtype=DT)                            # :32
    C0 = torch.ones((x1.shape[0], x2.shape[0]), dtype=DT)                            # :33
    x1s = x1 / llmb                                                                  # :35
    x2s = x2 / llmb                                                                  # :36
    for j in range(d):                                                               # :37-40
        S = torch.abs(x1s[:, j].reshape(-1, 1) - x2s[:, j])
        C0 = C0 * (1 + S)
        V = V - S
    C0 = C0 * torch.exp(V)                                                           # :42

    nug = lnug / (1 + lnug)                                                          # :45
    if x1.shape != x2.shape:                                                         # :46-47
        C = (1 - nug) * C0
    elif bool(torch.all(torch.eq(x1, x2))):                                          # :49-51
        C = (1 - nug) * C0 + nug * torch.eye(x1.shape[0], dtype=DT)
    else:                                                                            # :52-53
        C = (1 - nug) * C0
    return llmb0 * C                                                                 # :55


# ---------------------------------------------------------------------------
# lcgp.py: class LCGP
# ---------------------------------------------------------------------------
class LCGPOracle:
    """Restatement of `class LCGP` (lcgp.py:19-930) on torch.float64 CPU tensors."""

    def __init__(self, y=None, x=None, q: int = None, var_threshold: float = None,
                 diag_error_structure: list = None, parameter_clamp_flag: bool = False,
                 robust_mean: bool = True, submethod: str = 'full',
                 rep_standardize_ybar: bool = True, verbose: bool = False, skip_xnorm: bool = False):
        # skip_xnorm: do not evaluate the dead O(N^2 d) `xnorm` statistic of init_standard_x
        # (lcgp.py:304-309; unused by the model) -- only for the large bench configurations.
        self._skip_xnorm = skip_xnorm
        # lcgp.py:52-55
        self.verbose = verbose
        self.robust_mean = robust_mean
        self.rep_standardize_ybar = rep_standardize_ybar
        self.parameter_clamp_flag = parameter_clamp_flag
        # :60-61
        self.x = self._verify_data_types(x)
        self.y = self._verify_data_types(y)
        # :66-75
        self.method = 'LCGP'
        if submethod not in ['full', 'rep']:
            raise ValueError('Invalid submethod. Choices are \'full\' or \'rep\'.')
        self.submethod = submethod
        self.submethod_loss_map = {'full': self.neglpost, 'rep': self.neglpost_rep}
        self.submethod_predict_map = {'full': self.predict_full, 'rep': self.predict_rep}
        # :80-83
        if (q is not None) and (var_threshold is not None):
            raise ValueError('Include only q or var_threshold but not both.')
        self.q = q
        self.var_threshold = var_threshold
        # :88-92
        self.n, self.d, self.p = self.verify_dim(self.y, self.x)
        self.x_orig = self.x
        self.y_orig = self.y
        # :97
        self.x, self.x_min, self.x_max, _, self.xnorm = self.init_standard_x(self.x, skip_xnorm)
        self._rep_initialized = False

        if self.submethod == 'rep':                                    # :105-150
            xr, yr = self.x_orig.numpy(), self.y_orig.numpy()
            x_unique_np, inverse_np, counts_np = np.unique(
                xr, axis=0, return_inverse=True, return_counts=True)    # :353-355
            inverse_np = np.asarray(inverse_np).reshape(-1)
            n_unique = int(x_unique_np.shape[0])
            ybar_np = np.zeros((yr.shape[0], n_unique), dtype=np.float64)
            for i in range(n_unique):                                  # :364-366
                ybar_np[:, i] = yr[:, inverse_np == i].mean(axis=1)
            self.x_unique = torch.as_tensor(x_unique_np, dtype=DT)
            self.x_unique_s = (self.x_unique - self.x_min) / (self.x_max - self.x_min)   # :374
            self.group_ids = torch.as_tensor(inverse_np, dtype=torch.int32)
            self.r = torch.as_tensor(counts_np.astype(np.int32))
            self.R = torch.diag(self.r.to(DT))
            self.ybar = torch.as_tensor(ybar_np, dtype=DT)
            self.ybar_mean, self.ybar_std = self._compute_center_spread(self.ybar)     # :131
            self.ybar_s = (self.ybar - self.ybar_mean) / self.ybar_std
            self.n = n_unique
            self.d = int(xr.shape[1])
            self.p = int(yr.shape[0])
            self._rep_initialized = True
        else:                                                          # :155-156
            self.y, self.ymean, self.ystd, _ = self.init_standard_y(self.y)

        self.g, self.phi, self.diag_D, self.q = self.init_phi(var_threshold)          # :164
        self.Tks = None

        if diag_error_structure is None:                               # :171-176
            self.diag_error_structure = [1] * int(self.p)
        else:
            self.diag_error_structure = diag_error_structure
        assert sum(self.diag_error_structure) == self.y.shape[0], \
            'Sum of error_structure should equal the output dimension.'

        # bijectors, :181-211
        self.bij_lLmb = SoftClip(1e-6, 1e4)
        self.bij_lLmb0 = SoftClip(1e-4, 1e4)
        self.bij_lnug = SoftClip(math.exp(-16.0), math.exp(-2.0))
        self.init_params()                                             # :213

        self.CinvMs = None   # the reference fills q x n x n NaN placeholders (:218-222)
        self.Ths = None
        self.mks = None

    # -- utils ---------------------------------------------------------------
    @staticmethod
    def _verify_data_types(t):                                         # :248-258
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(np.asarray(t), dtype=DT)
        t = t.to(DT)
        if t.ndim < 2:
            t = t.unsqueeze(1)
        return t

    @staticmethod
    def verify_dim(y, x):                                              # :260-270
        p, ny = y.shape[0], y.shape[1]
        nx, d = x.shape[0], x.shape[1]
        assert ny == nx, 'Number of inputs (x) differs from number of outputs (y), y.shape[1] != x.shape[0]'
        return int(nx), int(d), int(p)

    def tx_x(self, xs):                                                # :280-284
        return xs * (self.x_max - self.x_min) + self.x_min

    def tx_y(self, ys):                                                # :286-290
        return ys * self.ystd + self.ymean

    @staticmethod
    def init_standard_x(x, skip_xnorm=False):                          # :295-310
        x_max = x.max(dim=0).values
        x_min = x.min(dim=0).values
        xs = (x - x_min) / (x_max - x_min)
        xnorm = torch.zeros(x.shape[1], dtype=DT)
        for j in range(0 if skip_xnorm else x.shape[1]):
            xdist = torch.abs(x[:, j].reshape(-1, 1) - x[:, j])
            xnorm[j] = xdist[xdist > 0].mean()
        return xs, x_min, x_max, x, xnorm

    def init_standard_y(self, y):                                      # :312-324
        if self.robust_mean:
            ycenter = percentile50_nearest(y)
            yspread = percentile50_nearest(torch.abs(y - ycenter))
        else:
            ycenter = y.mean(dim=1, keepdim=True)
            yspread = y.std(dim=1, keepdim=True, unbiased=False)
        return (y - ycenter) / yspread, ycenter, yspread, y

    def _compute_center_spread(self, Y):                               # :383-395
        if self.robust_mean:
            c = percentile50_nearest(Y)
            s = percentile50_nearest(torch.abs(Y - c))
        else:
            c = Y.mean(dim=1, keepdim=True)
            s = Y.std(dim=1, keepdim=True, unbiased=False)
        s = torch.where(s > 0, s, torch.ones_like(s))
        return c, s

    def _get_phi_input(self):                                          # :439-452
        if self.submethod != 'rep':
            return self.y
        if self.rep_standardize_ybar and hasattr(self, 'ybar_s'):
            return self.ybar_s
        return self.ybar

    def init_phi(self, var_threshold=None):                            # :454-485
        Y = self._get_phi_input()
        n, p = int(self.n), int(self.p)
        U, s, _ = torch.linalg.svd(Y, full_matrices=False)
        if (self.q is None) and (var_threshold is None):
            q = p
        elif (self.q is None) and (var_threshold is not None):
            cumvar = np.cumsum(s.numpy() ** 2) / np.sum(s.numpy() ** 2)
            q = int(np.argmax(cumvar > var_threshold) + 1) if np.any(cumvar > var_threshold) else p
        else:
            q = int(self.q)
        assert U.shape[1] == min(n, p)
        phi = U[:, :q] * math.sqrt(n) / s[:q]
        diag_D = (phi ** 2).sum(dim=0)
        g = phi.T @ Y
        return g, phi, diag_D, q

    # -- parameters ------------------------------------------------------------
    def init_params(self):                                             # :490-513
        x = self.x.numpy()
        d = int(self.d)
        llmb = np.exp(0.5 * np.log(d) + np.log(np.std(x, axis=0)))
        lLmb = np.tile(llmb, self.q).reshape((self.q, d))
        lLmb0 = np.ones(self.q)
        lnug = np.exp(-10.0) * np.ones(self.q)
        es = self.diag_error_structure
        lsig = np.zeros(len(es))
        col = 0
        ynp = self.y.numpy()
        for k in range(len(es)):
            lsig[k] = np.log(np.var(ynp[col:col + es[k]]))
            col += es[k]
        self.set_constrained(lLmb, lLmb0, lsig, lnug)

    def set_constrained(self, lLmb, lLmb0, lsigma2s, lnugGPs):
        """Assign constrained values; unconstrained variables are what the optimizer sees."""
        T = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64), dtype=DT)
        self.u_lLmb = self.bij_lLmb.inverse(T(lLmb)).detach().requires_grad_(True)
        self.u_lLmb0 = self.bij_lLmb0.inverse(T(lLmb0)).detach().requires_grad_(True)
        self.u_lnugGPs = self.bij_lnug.inverse(T(lnugGPs)).detach().requires_grad_(True)
        self.u_lsigma2s = T(lsigma2s).clone().detach().requires_grad_(True)

    @property
    def trainable_variables(self):
        # tf.Module attribute-name order: lLmb, lLmb0, lnugGPs, lsigma2s
        return [self.u_lLmb, self.u_lLmb0, self.u_lnugGPs, self.u_lsigma2s]

    @property
    def lLmb(self):
        return self.bij_lLmb.forward(self.u_lLmb)

    @property
    def lLmb0(self):
        return self.bij_lLmb0.forward(self.u_lLmb0)

    @property
    def lnugGPs(self):
        return self.bij_lnug.forward(self.u_lnugGPs)

    @property
    def lsigma2s(self):
        return self.u_lsigma2s

    def get_param(self):                                               # :515-532
        es = self.diag_error_structure
        reps = torch.as_tensor(es, dtype=torch.long)
        built = torch.repeat_interleave(self.lsigma2s, reps)
        return self.lLmb, self.lLmb0, built, self.lnugGPs

    # -- losses ------------------------------------------------------------------
    def loss(self):                                                    # :542-549
        try:
            return self.submethod_loss_map[self.submethod]()
        except KeyError:
            raise ValueError("Invalid submethod. Choices are 'full' or 'rep'.")

    def neglpost_rep(self, latents=None):                              # :554-630
        # `latents`: optional subset of k (bench.py times a bounded sample of the q-loop); None = all
        lLmb, lLmb0, lsigma2s, lnugGPs = self.get_param()
        xk = self.x_unique_s
        r = self.r.to(DT)
        n = float(self.n)
        p = float(self.p)
        D, phi = self.diag_D, self.phi

        sigma_var_raw = torch.exp(lsigma2s)                            # :572-574
        sigma_inv_sqrt_raw = torch.sqrt(1.0 / sigma_var_raw)
        if self.rep_standardize_ybar:                                  # :576-584
            ybar = self.ybar_s
            std = self.ybar_std[:, 0]
            sigma_var_used = sigma_var_raw / std ** 2
            sigma_inv_sqrt = sigma_inv_sqrt_raw * std
        else:
            ybar = self.ybar
            sigma_var_used = sigma_var_raw
            sigma_inv_sqrt = sigma_inv_sqrt_raw

        col_sq = ((ybar * sigma_inv_sqrt[:, None]) ** 2).sum(dim=0)    # :589-591
        nlp = 0.5 * (r * col_sq).sum()
        nlp = nlp + 0.5 * n * torch.log(sigma_var_used).sum()          # :594
        nlp = nlp - 0.5 * p * torch.log(r).sum()                       # :597
        sr = torch.sqrt(r)

        bkSb_sum = torch.zeros((), dtype=DT)
        logA_sum = torch.zeros((), dtype=DT)
        eye = torch.eye(int(self.n), dtype=DT)
        for k in (range(int(self.q)) if latents is None else latents):  # :605-624
            Ck = Matern32(xk, xk, llmb=lLmb[k], llmb0=lLmb0[k], lnug=lnugGPs[k])
            b_k = r * (ybar.T @ (sigma_inv_sqrt * phi[:, k]))          # :608-610
            d_k = D[k]
            Cb = Ck @ b_k                                              # :614
            A = eye + d_k * ((Ck * sr[None, :]) * sr[:, None])         # :616
            LA = torch.linalg.cholesky(A)                              # :617
            u = torch.sqrt(d_k) * (sr * Cb)                            # :618
            z = torch.cholesky_solve(u[:, None], LA)[:, 0]             # :619-620
            Sb = Cb - Ck @ (torch.sqrt(d_k) * (sr * z))                # :621
            bkSb_sum = bkSb_sum + b_k @ Sb                             # :623
            logA_sum = logA_sum + 2.0 * torch.log(torch.diagonal(LA)).sum()   # :624
        nlp = nlp - 0.5 * bkSb_sum + 0.5 * logA_sum                    # :626-627
        return nlp / n                                                 # :629

    def neglpost(self):                                                # :635-666 (eigh form, literal)
        lLmb, lLmb0, lsigma2s, lnugGPs = self.get_param()
        x, y = self.x, self.y
        n = float(self.n)
        D, phi = self.diag_D, self.phi
        psi_c = phi.T / torch.sqrt(torch.exp(lsigma2s))                # :646
        nlp = torch.zeros((), dtype=DT)
        for k in range(int(self.q)):                                   # :650-661
            Ck = Matern32(x, x, llmb=lLmb[k], llmb0=lLmb0[k], lnug=lnugGPs[k])
            Wk, Uk = torch.linalg.eigh(Ck)
            Qk = Uk @ (torch.diag(1 / (D[k] + 1 / Wk)) @ Uk.T)         # :654
            # :655-661  sum(yQk * (y^T Pk^T)^T) == (y^T psi)^T Qk (y^T psi); the reference forms the
            # p x p outer product Pk and a p x n GEMM; the value is the same quadratic form.
            Pk = psi_c[k][:, None] @ psi_c[k][None, :]
            yQk = y @ Qk
            yPk = y.T @ Pk.T
            nlp = nlp + 0.5 * torch.log(1 + D[k] * Wk).sum()
            nlp = nlp - 0.5 * (yQk * yPk.T).sum()
        nlp = nlp + n / 2 * lsigma2s.sum()                             # :663
        nlp = nlp + 0.5 * ((y.T / torch.sqrt(torch.exp(lsigma2s))) ** 2).sum()   # :664
        return nlp

    def neglpost_chol(self, latents=None):
        """Full-mode objective in single-Cholesky form: algebraically identical to
        `neglpost` (eigh form) but O(n^3/3) per latent; used as the oracle at sizes where eigh +
        autograd is too slow.  Checked against `neglpost` in tests/test_oracle_selfcheck.py."""
        lLmb, lLmb0, lsigma2s, lnugGPs = self.get_param()
        x, y = self.x, self.y
        n = float(self.n)
        D, phi = self.diag_D, self.phi
        sinv = torch.exp(-0.5 * lsigma2s)
        eye = torch.eye(int(self.n), dtype=DT)
        nlp = torch.zeros((), dtype=DT)
        for k in (range(int(self.q)) if latents is None else latents):
            Ck = Matern32(x, x, llmb=lLmb[k], llmb0=lLmb0[k], lnug=lnugGPs[k])
            b = y.T @ (sinv * phi[:, k])
            LA = torch.linalg.cholesky(eye + D[k] * Ck)
            alpha = torch.cholesky_solve(b[:, None], LA)[:, 0]
            nlp = nlp + torch.log(torch.diagonal(LA)).sum() - 0.5 * (b @ (Ck @ alpha))
        nlp = nlp + n / 2 * lsigma2s.sum()
        nlp = nlp + 0.5 * ((y.T * sinv) ** 2).sum()
        return nlp

    # -- flat-vector interface for the optimizer ---------------------------------
    def _flat_get(self):
        return torch.cat([v.detach().reshape(-1) for v in self.trainable_variables]).numpy().copy()

    def _flat_set(self, vec):
        vec = torch.as_tensor(np.asarray(vec, dtype=np.float64), dtype=DT)
        o = 0
        for v in self.trainable_variables:
            m = v.numel()
            v.data = vec[o:o + m].reshape(v.shape).clone()
            o += m

    def loss_and_grad(self, loss_fn=None):
        """Returns (loss float, flat gradient wrt unconstrained variables) by autograd."""
        for v in self.trainable_variables:
            v.grad = None
        val = (loss_fn or self.loss)()
        val.backward()
        g = torch.cat([v.grad.reshape(-1) for v in self.trainable_variables]).numpy().copy()
        return float(val.detach()), g

    def grad_constrained(self, loss_fn=None):
        """Gradient wrt the constrained parameter values (lLmb, lLmb0, lsigma2s(G), lnugGPs)."""
        lLmb = self.lLmb.detach().requires_grad_(True)
        lLmb0 = self.lLmb0.detach().requires_grad_(True)
        lsig = self.lsigma2s.detach().requires_grad_(True)
        lnug = self.lnugGPs.detach().requires_grad_(True)
        saved = self.get_param

        def gp():
            reps = torch.as_tensor(self.diag_error_structure, dtype=torch.long)
            return lLmb, lLmb0, torch.repeat_interleave(lsig, reps), lnug
        self.get_param = gp
        try:
            val = (loss_fn or self.loss)()
            val.backward()
        finally:
            self.get_param = saved
        return float(val.detach()), lLmb.grad, lLmb0.grad, lsig.grad, lnug.grad

    def fit(self, verbose=False, method='L-BFGS-B', loss_fn=None, **opts):   # :537-540
        """gpflow.optimizers.Scipy().minimize(self.loss, self.trainable_variables, compile=False)
        == scipy.optimize.minimize(method='L-BFGS-B', jac=True) on the unconstrained variables."""
        self.CinvMs = self.Ths = self.Tks = self.mks = None
        if method == 'L-BFGS-B':
            import scipy.optimize

            def fun(v):
                self._flat_set(v)
                return self.loss_and_grad(loss_fn)
            res = scipy.optimize.minimize(fun, self._flat_get(), jac=True, method='L-BFGS-B',
                                          options=opts or None)
            self._flat_set(res.x)
            self.opt_result = res
            return res
        elif method == 'torch-lbfgs':
            kw = dict(lr=1.0, max_iter=opts.pop('max_iter', 500), history_size=10,
                      line_search_fn='strong_wolfe', tolerance_grad=1e-9, tolerance_change=1e-12)
            kw.update(opts)
            opt = torch.optim.LBFGS(self.trainable_variables, **kw)
            nev = [0]

            def closure():
                opt.zero_grad()
                val = (loss_fn or self.loss)()
                val.backward()
                nev[0] += 1
                return val
            opt.step(closure)
            self.opt_result = {'nfev': nev[0]}
            return self.opt_result
        raise ValueError(method)

    # -- aux predictive quantities -----------------------------------------------
    @torch.no_grad()
    def compute_aux_predictive_quantities(self):                       # :685-726
        if hasattr(self, 'x_unique') and hasattr(self, 'ybar'):
            self._compute_aux_predictive_quantities_rep()
            return
        x = self.x
        lLmb, lLmb0, lsigma2s, lnugGPs = self.get_param()
        D = self.diag_D
        B = (self.y.T / torch.sqrt(torch.exp(lsigma2s))) @ self.phi   # :697
        q, n = int(self.q), int(self.n)
        CinvM = torch.zeros((q, n), dtype=DT)
        Th = torch.zeros((q, n, n), dtype=DT)
        for k in range(q):                                             # :702-716
            Ck = Matern32(x, x, llmb=lLmb[k], llmb0=lLmb0[k], lnug=lnugGPs[k])
            Wk, Uk = torch.linalg.eigh(Ck)
            IpdkCkinv = Uk @ (torch.diag(1.0 / (1.0 + D[k] * Wk)) @ Uk.T)
            CinvM[k] = IpdkCkinv @ B.T[k]
            Th[k] = Uk @ (torch.diag(torch.sqrt((D[k] * Wk ** 2) / (Wk ** 2 + D[k] * Wk ** 3))) @ Uk.T)
        self.CinvMs = CinvM
        self.Ths = Th

    @torch.no_grad()
    def _compute_aux_predictive_quantities_rep(self, stable: bool = True):   # :728-803
        """stable=False follows lcgp.py:783-788 literally (explicit inv(C), ~1e-9 abs error at
        n=60); stable=True uses the algebraically identical
        Tk = d_k (sqrt r sqrt r^T) o A_k^{-1}, which is the form predictions are judged against."""
        lLmb, lLmb0, lsigma2s, lnugGPs = self.get_param()
        xk = self.x_unique_s
        r = self.r.to(DT)
        R = self.R
        D, phi = self.diag_D, self.phi
        use_std = self.rep_standardize_ybar
        ybar = self.ybar_s if use_std else self.ybar
        sigma_inv_sqrt_used = torch.exp(-0.5 * lsigma2s)              # :746-751
        if use_std:
            sigma_inv_sqrt_used = sigma_inv_sqrt_used * self.ybar_std[:, 0]
        q, n, p = int(self.q), int(self.n), int(self.p)
        if q == p or p == 1:                                           # :754 as coded (broadcast q==p)
            self.psi_c = phi.T / sigma_inv_sqrt_used[:, None]
        else:                                                          # reference raises here (Appendix B-3)
            self.psi_c = phi.T / sigma_inv_sqrt_used[None, :]
        CinvM = torch.zeros((q, n), dtype=DT)
        Tks = torch.zeros((q, n, n), dtype=DT)
        mks = torch.zeros((q, n), dtype=DT)
        sr = torch.sqrt(r)
        eye = torch.eye(n, dtype=DT)
        for k in range(q):                                             # :765-790
            Ck = Matern32(xk, xk, llmb=lLmb[k], llmb0=lLmb0[k], lnug=lnugGPs[k])
            b_k = r * (ybar.T @ (sigma_inv_sqrt_used * phi[:, k]))
            d_k = D[k]
            Cb = Ck @ b_k
            A = eye + d_k * ((Ck * sr[None, :]) * sr[:, None])
            LA = torch.linalg.cholesky(A)
            u = torch.sqrt(d_k) * (sr * Cb)
            z = torch.cholesky_solve(u[:, None], LA)[:, 0]
            m_k = Cb - Ck @ (torch.sqrt(d_k) * (sr * z))
            CinvM[k] = b_k - d_k * (R @ m_k)                            # :781
            if stable:
                Ainv = torch.cholesky_inverse(LA)
                Tks[k] = d_k * (sr[:, None] * sr[None, :]) * Ainv
            else:                                                      # :783-788
                LC = torch.linalg.cholesky(Ck)
                invC = torch.cholesky_solve(eye, LC)
                V_k = torch.linalg.inv(invC + d_k * R)
                Tks[k] = invC - invC @ V_k @ invC
            mks[k] = m_k
        self.mks, self.CinvMs, self.Tks, self.Ths = mks, CinvM, Tks, None

    # -- prediction -----------------------------------------------------------------
    def predict(self, x0, return_fullcov=False):                       # :671-680
        x0 = self._verify_data_types(x0)
        try:
            call = self.submethod_predict_map[self.submethod]
        except KeyError:
            raise KeyError('Invalid submethod.  Choices are \'full\' or \'rep\'.')
        with torch.no_grad():
            res = call(x0=x0, return_fullcov=return_fullcov)
        return tuple(t.detach() if t is not None else None for t in res)

    def predict_full(self, x0, return_fullcov=False):                  # :808-859
        if self.CinvMs is None or self.Ths is None:
            self.compute_aux_predictive_quantities()
        x = self.x
        lLmb, lLmb0, lsigma2s, lnugGPs = self.get_param()
        phi, CinvM, Th = self.phi, self.CinvMs, self.Ths
        x0 = (x0 - self.x_min) / (self.x_max - self.x_min)             # :822
        n0 = x0.shape[0]
        q = int(self.q)
        ghat = torch.zeros((q, n0), dtype=DT)
        gvar = torch.zeros((q, n0), dtype=DT)
        for k in range(q):                                             # :827-835
            c00k = Matern32(x0, x0, llmb=lLmb[k], llmb0=lLmb0[k], lnug=lnugGPs[k], diag_only=True)
            c0k = Matern32(x0, x, llmb=lLmb[k], llmb0=lLmb0[k], lnug=lnugGPs[k])
            ghat[k] = c0k @ CinvM[k]
            gvar[k] = c00k - ((c0k @ Th[k]) ** 2).sum(dim=1)
        self.ghat, self.gvar = ghat, gvar
        psi = phi.T * torch.sqrt(torch.exp(lsigma2s))                  # :840
        predmean = psi.T @ ghat
        confvar = gvar.T @ psi ** 2
        predvar = confvar + torch.exp(lsigma2s)
        ypred = self.tx_y(predmean)
        yconfvar = confvar.T * self.ystd ** 2
        ypredvar = predvar.T * self.ystd ** 2
        if return_fullcov:                                             # :850-857
            CH = torch.einsum('kn,kp->npk', torch.sqrt(gvar), psi)
            full = CH @ CH.transpose(1, 2)
            full = full + torch.diag(torch.exp(lsigma2s))[None]
            sv = self.ystd[:, 0]
            full = full * (sv[:, None] * sv[None, :])[None]
            return ypred, ypredvar, yconfvar, full
        return ypred, ypredvar, yconfvar

    def predict_rep(self, x0, return_fullcov=False):                   # :864-930
        if self.Tks is None or self.CinvMs is None:
            self._compute_aux_predictive_quantities_rep()
        lLmb, lLmb0, lsigma2s, lnugGPs = self.get_param()
        phi = self.phi
        Xtrain, Tks, CinvM = self.x_unique_s, self.Tks, self.CinvMs
        x0 = (x0 - self.x_min) / (self.x_max - self.x_min)             # :877
        n0 = x0.shape[0]
        q = int(self.q)
        ghat = torch.zeros((q, n0), dtype=DT)
        gvar = torch.zeros((q, n0), dtype=DT)
        for k in range(q):                                             # :883-897
            c00k = Matern32(x0, x0, llmb=lLmb[k], llmb0=lLmb0[k], lnug=lnugGPs[k], diag_only=True)
            c0k = Matern32(x0, Xtrain, llmb=lLmb[k], llmb0=lLmb0[k], lnug=lnugGPs[k])
            ghat[k] = c0k @ CinvM[k]
            gvar[k] = c00k - ((c0k @ Tks[k]) * c0k).sum(dim=1)
        self.ghat, self.gvar = ghat, gvar
        sigma_var_raw = torch.exp(lsigma2s)                            # :904-913
        sigma_sqrt_raw = torch.sqrt(sigma_var_raw)
        if self.rep_standardize_ybar:
            std = self.ybar_std[:, 0]
            sigma_sqrt_used = sigma_sqrt_raw / std
            sigma_var_used = sigma_var_raw / std ** 2
        else:
            sigma_sqrt_used, sigma_var_used = sigma_sqrt_raw, sigma_var_raw
        Psi = phi * sigma_sqrt_used[:, None]                           # :915
        predmean_used = Psi @ ghat
        confvar_used = (Psi ** 2) @ gvar
        predvar_used = confvar_used + sigma_var_used[:, None]
        if self.rep_standardize_ybar:                                  # :921-926
            ypred = predmean_used * self.ybar_std + self.ybar_mean
            yconfvar = confvar_used * self.ybar_std ** 2
            ypredvar = predvar_used * self.ybar_std ** 2
        else:
            ypred, yconfvar, ypredvar = predmean_used, confvar_used, predvar_used
        if return_fullcov:
            return ypred, ypredvar, yconfvar, None
        return ypred, ypredvar, yconfvar


# ---------------------------------------------------------------------------
# evaluation.py:5-63 (metrics used to replay the notebook's golden numbers)
# ---------------------------------------------------------------------------
def rmse(y, ypredmean):                                                # evaluation.py:5-9
    return float(np.sqrt(np.mean((y - ypredmean) ** 2)))


def normalized_rmse(y, ypredmean):                                     # evaluation.py:12-18
    rng = (np.max(y, axis=1) - np.min(y, axis=1)).reshape(y.shape[0], 1)
    return float(np.sqrt(np.mean(((y - ypredmean) / rng) ** 2)))


def dss_diag(y, ypredmean, ypredvar):                                  # evaluation.py:21-49, use_diag=True
    n = y.shape[1]
    return float((np.log(ypredvar).sum() + ((y - ypredmean) ** 2 / ypredvar).sum()) / n)


def intervalstats(y, ypredmean, ypredvar):                             # evaluation.py:52-63
    import scipy.stats as sps
    lo = ypredmean + np.sqrt(ypredvar) * sps.norm.ppf(0.025)
    hi = ypredmean + np.sqrt(ypredvar) * sps.norm.ppf(0.975)
    return float(np.mean(np.logical_and(y <= hi, y >= lo))), float(np.mean(hi - lo))
