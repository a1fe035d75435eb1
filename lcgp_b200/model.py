"""`LCGP` -- drop-in for the reference's model object (src/lcgp/lcgp.py:19-930) with the
emulator-fitting hot path on sm_100a CUDA kernels behind the C-ABI of include/lcgp_b200.h.

What stays on the host (torch CPU, float64): input checking, standardisation, replicate grouping,
the SVD basis, the parameter container with its soft-clip bijectors, and the L-BFGS outer loop.
What runs on the GPU, once per objective evaluation: kernel-matrix build, batched Cholesky,
triangular inverse, solves, log-determinants and the analytic gradient (one `lcgp_nll_grad_host`
call), and -- for `predict` -- the cross-covariance products against the same factor.

There is no CPU fallback: `loss()`, `fit()` and `predict()` raise without a CUDA device / built
library.  With `torch.distributed` initialised (one process per GPU) the q independent latent GPs
are sharded round-robin over the ranks and one small all-reduce per evaluation combines the
objective and the gradient; every rank then holds identical values, so the
replicated host optimizer stays in lock-step.
"""
from __future__ import annotations

import math
import os
from typing import Optional

import numpy as np
import torch

from . import _cabi
from .parameter import Parameter, SoftClip

DT = torch.float64


# --------------------------------------------------------------------------------------------
# small host helpers
# --------------------------------------------------------------------------------------------
def _as_tensor2d(t):
    """lcgp.py:248-258: cast to float64 tensor, at least 2-D (a 1-D input becomes a column)."""
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(np.asarray(t), dtype=DT)
    t = t.detach().to('cpu', DT)
    if t.ndim < 2:
        t = t.unsqueeze(1)
    return t


def _nearest_rank_median(Y):
    """50th percentile along axis 1 with 'nearest' interpolation (what tfp.stats.percentile does by
    default, lcgp.py:317-318/388-389): element round((m-1)/2) (half to even) of the ascending sort."""
    m = Y.shape[1]
    k = int(np.round((m - 1) * 0.5))
    return torch.kthvalue(Y, k + 1, dim=1, keepdim=True).values


def _itensor(v):
    return torch.tensor(int(v), dtype=torch.int32)


class _NegLogPost(torch.autograd.Function):
    """Objective of all latents as one autograd node: forward runs the CUDA path and keeps the
    analytic gradient; backward scales it.  The soft-clip chain rule and the
    diag_error_structure segment-sum are left to torch autograd on the host."""

    @staticmethod
    def forward(ctx, model, lLmb, lLmb0, lsig_p, lnug):
        need_grad = any(ctx.needs_input_grad[1:])
        val, grads = model._evaluate(lLmb.detach(), lLmb0.detach(), lsig_p.detach(), lnug.detach(), need_grad)
        if need_grad:
            ctx.save_for_backward(*grads)
        return val

    @staticmethod
    def backward(ctx, gout):
        g_lLmb, g_lLmb0, g_lsig, g_lnug = ctx.saved_tensors
        return None, gout * g_lLmb, gout * g_lLmb0, gout * g_lsig, gout * g_lnug


class CudaEngine:
    """Device-resident constant data of this rank's latents + the C-ABI calls on them."""

    def __init__(self, n, d, p, X, sr, YR, w, t, phi_loc, D_loc, scale, sum_log_r, include_host_terms, device=None,
                 stream_groups=0):
        _cabi.require_cuda()
        self.group_flags = (int(stream_groups) & 15) << 4   # 0 = library default (see include/lcgp_b200.h)
        self.lib = _cabi.lib()
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.n, self.d, self.p = int(n), int(d), int(p)
        self.q_loc = int(phi_loc.shape[1])
        if self.d > _cabi.MAX_D:
            raise ValueError(f'lcgp_b200 supports input dimension d <= {_cabi.MAX_D}')
        dev = self.device
        c = lambda a: a.to(dev, DT).contiguous()
        self.X, self.sr, self.YR, self.w, self.t = c(X), c(sr), c(YR), c(w), c(t)
        self.phi, self.D = c(phi_loc), c(D_loc)
        self.prob = _cabi.Problem(n=self.n, d=self.d, p=self.p, q_loc=self.q_loc,
                                  include_host_terms=int(bool(include_host_terms)), n_emu=0, emu_consts=None,
                                  scale=float(scale), sum_log_r=float(sum_log_r),
                                  X=self.X.data_ptr(), sr=self.sr.data_ptr(), YR=self.YR.data_ptr(),
                                  w=self.w.data_ptr(), t=self.t.data_ptr(), phi=self.phi.data_ptr(),
                                  D=self.D.data_ptr())
        self.ws_bytes = int(self.lib.lcgp_workspace_bytes(self.n, self.d, self.p, self.q_loc))
        self.ws = torch.empty(self.ws_bytes // 8, dtype=DT, device=dev)
        self.out_len = int(self.lib.lcgp_out_len(self.p, self.d, self.q_loc))
        q, dd = self.q_loc, self.d
        # pinned host staging: [lLmb | lLmb0 | lnug | lsig_p] and the result vector
        # (two pinned buffers, used alternately: evaluate_device does not synchronise the host, so the copy of the
        # previous call may still be pending when the next call stages its parameters)
        self._h_pars = [torch.empty(q * dd + 2 * q + self.p, dtype=DT).pin_memory() for _ in range(2)]
        self._h_par_free = [None, None]      # event recorded after the last asynchronous H2D copy out of each buffer
        self._h_cur = 0
        self.h_par = self._h_pars[0]
        self.h_out = torch.empty(self.out_len, dtype=DT).pin_memory()
        self.h_info = torch.zeros(max(q, 1), dtype=torch.int32).pin_memory()
        self.d_par = torch.empty(q * dd + 2 * q + self.p, dtype=DT, device=dev)
        self.d_out = torch.empty(self.out_len, dtype=DT, device=dev)
        self.d_info = torch.zeros(max(q, 1), dtype=torch.int32, device=dev)
        self.h2d_bytes = self.h_par.numel() * 8
        self.d2h_bytes = self.out_len * 8 + q * 4
        # launch-bound sizes (n <= 2048): evaluations replay a captured CUDA graph (lcgp_plan_*)
        self.use_plans = _cabi.padded(self.n) // _cabi.NB <= 16 and os.environ.get('LCGP_GRAPHS', '1') != '0'
        self._plans = {}
        self._scratch = None

    def _stage(self, lLmb, lLmb0, lnug, lsig_p, rotate=False):
        q, dd = self.q_loc, self.d
        if rotate:      # asynchronous callers: take the other pinned buffer, once its last copy has left it
            self._h_cur ^= 1
            ev = self._h_par_free[self._h_cur]
            if ev is not None:
                ev.synchronize()
        h = self._h_pars[self._h_cur] if rotate else self.h_par
        h[:q * dd].copy_(lLmb.reshape(-1))
        h[q * dd:q * dd + q].copy_(lLmb0.reshape(-1))
        h[q * dd + q:q * dd + 2 * q].copy_(lnug.reshape(-1))
        h[q * dd + 2 * q:].copy_(lsig_p.reshape(-1))
        return q * dd, q * dd + q, q * dd + 2 * q

    def _upload_async(self):
        """d_par <- the pinned buffer _stage(rotate=True) just filled, without a host synchronisation."""
        self.d_par.copy_(self._h_pars[self._h_cur], non_blocking=True)
        ev = self._h_par_free[self._h_cur]
        if ev is None:
            ev = self._h_par_free[self._h_cur] = torch.cuda.Event()
        ev.record()

    def _events_arg(self, events):
        if events is None:
            return None
        import ctypes as C
        arr = (C.c_void_p * _cabi.N_STAGE_EVENTS)(*[int(e.cuda_event) for e in events])
        return arr

    def evaluate(self, lLmb, lLmb0, lnug, lsig_p, with_grad=True, events=None):
        """Host parameters in -> `out` vector (CPU tensor, layout of lcgp_out_len) via one
        lcgp_nll_grad_host call (H2D of the parameters, all kernels, D2H of the result, sync)."""
        o1, o2, o3 = self._stage(lLmb, lLmb0, lnug, lsig_p)
        base = self.h_par.data_ptr()
        if self.use_plans and events is None:
            self._run_plan(int(bool(with_grad)) | self.group_flags)
            self._check_info(self.h_info)
            return self.h_out.clone()
        with torch.cuda.device(self.device):
            rc = self.lib.lcgp_nll_grad_host(self.prob, base, base + 8 * o1, base + 8 * o2, base + 8 * o3,
                                             self.ws.data_ptr(), self.ws_bytes, self.h_out.data_ptr(),
                                             self.h_info.data_ptr(), int(bool(with_grad)) | self.group_flags,
                                             self._events_arg(events), _cabi.stream_ptr())
        _cabi.check(rc, 'lcgp_nll_grad_host')
        self._check_info(self.h_info)
        return self.h_out.clone()

    def _run_plan(self, flags):
        import ctypes as C
        plan = self._plans.get(flags)
        with torch.cuda.device(self.device):
            if plan is None:
                h = C.c_void_p()
                rc = self.lib.lcgp_plan_create(self.prob, self.ws.data_ptr(), self.ws_bytes, self.h_par.data_ptr(),
                                               self.h_out.data_ptr(), self.h_info.data_ptr(), flags, C.byref(h))
                _cabi.check(rc, 'lcgp_plan_create')
                plan = self._plans[flags] = h
            rc = self.lib.lcgp_plan_run(plan, _cabi.stream_ptr())   # ordered after the current stream; synchronises
        _cabi.check(rc, 'lcgp_plan_run')

    def plan_is_graph(self):
        """True when every plan created so far replays a captured graph (False: none yet, or capture refused)."""
        return bool(self._plans) and all(self.lib.lcgp_plan_is_graph(h) for h in self._plans.values())

    def __del__(self):
        try:
            for h in getattr(self, '_plans', {}).values():
                self.lib.lcgp_plan_destroy(h)
        except Exception:
            pass

    def evaluate_device(self, lLmb, lLmb0, lnug, lsig_p, with_grad=True, events=None):
        """Same, but the result stays on the device (multi-rank path: all-reduce follows on the
        same stream).  No host synchronisation."""
        o1, o2, o3 = self._stage(lLmb, lLmb0, lnug, lsig_p, rotate=True)
        with torch.cuda.device(self.device):
            self._upload_async()
            base = self.d_par.data_ptr()
            rc = self.lib.lcgp_nll_grad(self.prob, base, base + 8 * o1, base + 8 * o2, base + 8 * o3,
                                        self.ws.data_ptr(), self.ws_bytes, self.d_out.data_ptr(),
                                        self.d_info.data_ptr(), int(bool(with_grad)) | self.group_flags,
                                        self._events_arg(events), _cabi.stream_ptr())
        _cabi.check(rc, 'lcgp_nll_grad')
        return self.d_out

    def pack_sharded(self, loc_of, q, flat):
        """d_out / d_info of the last evaluate_device -> the flat all-reduce vector (lcgp_pack_sharded), in place."""
        with torch.cuda.device(self.device):
            rc = self.lib.lcgp_pack_sharded(self.d_out.data_ptr(), self.d_info.data_ptr(), loc_of.data_ptr(), self.p,
                                            self.d, int(q), self.q_loc, flat.data_ptr(), _cabi.stream_ptr())
        _cabi.check(rc, 'lcgp_pack_sharded')

    def _check_info(self, info):
        bad = torch.nonzero(info[:self.q_loc])
        if bad.numel():
            k = int(bad[0])
            raise RuntimeError(f'lcgp_b200: Cholesky failed for local latent {k} at pivot {int(info[k])} '
                               f'(A_k = I + d_k R^1/2 C_k R^1/2 has eigenvalues >= 1, so this means NaN/inf inputs)')

    def predict_latents(self, lLmb, lLmb0, lnug, x0s, same):
        """ghat, gvar (q_loc x n0) on the device from the factor left by the last evaluate()."""
        o1, o2, _ = self._stage(lLmb, lLmb0, lnug, torch.zeros(self.p, dtype=DT), rotate=True)
        n0 = int(x0s.shape[0])
        with torch.cuda.device(self.device):
            self._upload_async()
            x0d = x0s.to(self.device, DT).contiguous()
            need = int(self.lib.lcgp_predict_scratch_bytes(self.n, self.q_loc, n0))
            if self._scratch is None or self._scratch.numel() * 8 < need:
                self._scratch = torch.empty(need // 8, dtype=DT, device=self.device)
            ghat = torch.empty((self.q_loc, n0), dtype=DT, device=self.device)
            gvar = torch.empty((self.q_loc, n0), dtype=DT, device=self.device)
            base = self.d_par.data_ptr()
            rc = self.lib.lcgp_predict(self.prob, base, base + 8 * o1, base + 8 * o2, self.ws.data_ptr(), self.ws_bytes,
                                       x0d.data_ptr(), n0, int(bool(same)), self._scratch.data_ptr(),
                                       self._scratch.numel() * 8, ghat.data_ptr(), gvar.data_ptr(), _cabi.stream_ptr())
        _cabi.check(rc, 'lcgp_predict')
        return ghat, gvar

    def grad_phi(self, lsig_p):
        """d objective / d phi[:, local latents] (p x q_loc, device) from the state the last evaluate(with_grad=True) left."""
        with torch.cuda.device(self.device):
            ls = lsig_p.to(self.device, DT).contiguous()
            g = torch.empty((self.p, self.q_loc), dtype=DT, device=self.device)
            rc = self.lib.lcgp_grad_phi(self.prob, ls.data_ptr(), self.ws.data_ptr(), self.ws_bytes, g.data_ptr(),
                                        _cabi.stream_ptr())
        _cabi.check(rc, 'lcgp_grad_phi')
        return g

    def aux(self):
        a = torch.empty((self.q_loc, self.n), dtype=DT, device=self.device)
        m = torch.empty((self.q_loc, self.n), dtype=DT, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.lcgp_get_aux(self.prob, self.ws.data_ptr(), self.ws_bytes, a.data_ptr(), m.data_ptr(),
                                       _cabi.stream_ptr())
        _cabi.check(rc, 'lcgp_get_aux')
        return a, m

    def ainv(self, k):
        A = torch.empty((self.n, self.n), dtype=DT, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.lcgp_get_Ainv(self.prob, self.ws.data_ptr(), self.ws_bytes, int(k), A.data_ptr(),
                                        _cabi.stream_ptr())
        _cabi.check(rc, 'lcgp_get_Ainv')
        return A


class LCGP:
    """Latent Component Gaussian Process.

    submethod='full' uses all observations (x, y); submethod='rep' groups replicated x rows and
    works on (x_unique, ybar).  Constructor arguments, attributes, methods and raised exception
    types follow lcgp.py:31-41 and the reference test-suite; tensors are torch.float64.
    Extra keyword arguments (not in the reference): `device`, `engine_factory` (tests only),
    `shard` (shard latents over torch.distributed ranks when a process group is initialised),
    `stream_groups` (internal CUDA stream groups per evaluation; 0 = library default).
    """

    def __init__(self, y=None, x=None, q: int = None, var_threshold: float = None,
                 diag_error_structure: list = None, parameter_clamp_flag: bool = False,
                 robust_mean: bool = True, submethod: str = 'full', rep_standardize_ybar: bool = True,
                 verbose: bool = False, device=None, shard: bool = True, engine_factory=None, stream_groups: int = 0,
                 device_preprocess: Optional[bool] = None):
        self.verbose = verbose
        self.robust_mean = robust_mean
        self._rep_standardize_ybar = bool(rep_standardize_ybar)
        self.parameter_clamp_flag = parameter_clamp_flag
        self._device = device
        self._engine_factory = engine_factory
        self._engine = None
        self._stream_groups = stream_groups
        # O(p N) preprocessing (replicate means, median / MAD, standardisation) on the device (csrc/prep.cu);
        # None = wherever a CUDA device and the library are present.  The host implementation is the one
        # SURVEY 8 a11 designates; both give the same bits except for the row sums w (reduction order).
        if device_preprocess is None:
            device_preprocess = engine_factory is None and torch.cuda.is_available() and _cabi.available()
        elif device_preprocess:
            _cabi.require_cuda()
        self._dev_prep = bool(device_preprocess)
        # sharding of the latents over ranks
        self._world, self._rank = 1, 0
        if shard and torch.distributed.is_available() and torch.distributed.is_initialized():
            self._world = torch.distributed.get_world_size()
            self._rank = torch.distributed.get_rank()

        self.x = _as_tensor2d(x)
        self.y = _as_tensor2d(y)

        self.method = 'LCGP'
        if submethod not in ['full', 'rep']:
            raise ValueError('Invalid submethod. Choices are \'full\' or \'rep\'.')
        self.submethod = submethod
        self.submethod_loss_map = {'full': self.neglpost, 'rep': self.neglpost_rep}
        self.submethod_predict_map = {'full': self.predict_full, 'rep': self.predict_rep}

        if (q is not None) and (var_threshold is not None):
            raise ValueError('Include only q or var_threshold but not both.')
        self.q = q
        self.var_threshold = var_threshold

        self.n, self.d, self.p = self.verify_dim(self.y, self.x)
        self.x_orig = self.x
        self.y_orig = self.y

        self.x, self.x_min, self.x_max, _ = self._standardize_x(self.x)
        self._xnorm = None
        self._rep_initialized = False

        if self.submethod == 'rep':
            (self.x_unique, self.x_unique_s, self.group_ids, self.r, _R, self.ybar, self.ybar_s,
             self.ybar_mean, self.ybar_std, self.n, self.d, self.p) = self.preprocess()
            self._rep_initialized = True
        else:
            self.y, self.ymean, self.ystd, _ = self.init_standard_y(self.y)

        self.g, self.phi, self.diag_D, self.q = self.init_phi(var_threshold=var_threshold)

        if diag_error_structure is None:
            self.diag_error_structure = [1] * int(self.p)
        else:
            self.diag_error_structure = diag_error_structure
        self.verify_error_structure(self.diag_error_structure, self.y)

        d = int(self.d)
        self.lLmb = Parameter(np.ones((self.q, d)), SoftClip(1e-6, 1e4), name='Latent GP log-scale')
        self.lLmb0 = Parameter(np.ones(self.q), SoftClip(1e-4, 1e4), name='Latent GP log-lengthscale')
        self.lsigma2s = Parameter(np.ones(len(self.diag_error_structure)), None, name='Diagonal error log-variance')
        self.lnugGPs = Parameter(np.ones(self.q) * 1e-6, SoftClip(math.exp(-16.0), math.exp(-2.0)),
                                 name='Latent GP nugget scale')
        self.init_params()

        self._local_idx = torch.arange(self.q)[self._rank::self._world]   # empty when this rank owns no latent

        self._invalidate_aux()
        self._ghat = self._gvar = None          # latent predictive moments of the last predict(): CPU copies made on demand
        self._ghat_dev = self._gvar_dev = None
        self.psi_c = None
        self.n_evals = 0

    @property
    def rep_standardize_ybar(self):
        return self._rep_standardize_ybar

    @rep_standardize_ybar.setter
    def rep_standardize_ybar(self, flag):
        """The reference marks this flag "can toggle" (lcgp.py:49); the engine's constant data (ybar or ybar_s, t) depend
        on it, so a change drops the engine and the cached predictive quantities -- objective, factor and the output maps
        of predict_rep then all use the same standardisation.  phi / diag_D stay those of construction, as in the
        reference."""
        flag = bool(flag)
        if flag != self._rep_standardize_ybar:
            self._rep_standardize_ybar = flag
            self._engine = None
            if isinstance(getattr(self, 'q', None), int) and hasattr(self, 'CinvMs'):
                self._invalidate_aux()

    # ------------------------------------------------------------------ display
    def __repr__(self):
        rows = '\n'.join(f'\t\t{prm.name}\t{tuple(prm.shape)}\t{np.array2string(prm.numpy(), precision=4, threshold=8)}'
                         for prm in (self.lLmb, self.lLmb0, self.lsigma2s, self.lnugGPs))
        return ('LCGP(\n\tsubmethod:\t{:s}\n\toutput dimension:\t{:d}\n\tnumber of latent components:\t{:d}\n'
                '\tparameter_clamping:\t{:s}\n\trobust_standardization:\t{:s}\n'
                '\tdiagonal_error structure:\t{:s}\n\tparameters:\t\n{}\n)').format(
                    self.submethod, int(self.p), int(self.q), str(self.parameter_clamp_flag),
                    str(self.robust_mean), str(self.diag_error_structure), rows)

    # ------------------------------------------------------------------ checks / transforms
    @staticmethod
    def _verify_data_types(t):
        return _as_tensor2d(t)

    def verify_dim(self, y, x):
        """lcgp.py:260-270."""
        p, ny = y.shape[0], y.shape[1]
        nx, d = x.shape[0], x.shape[1]
        assert ny == nx, 'Number of inputs (x) differs from number of outputs (y), y.shape[1] != x.shape[0]'
        return _itensor(nx), _itensor(d), _itensor(p)

    @staticmethod
    def verify_error_structure(diag_error_structure, y):
        """lcgp.py:272-278."""
        assert sum(diag_error_structure) == y.shape[0], \
            'Sum of error_structure should equal the output dimension.'

    def tx_x(self, xs):
        return xs * (self.x_max - self.x_min) + self.x_min

    def tx_y(self, ys):
        return ys * self.ystd + self.ymean

    # ------------------------------------------------------------------ standardisation
    @staticmethod
    def _standardize_x(x):
        x_max = x.max(dim=0).values
        x_min = x.min(dim=0).values
        return (x - x_min) / (x_max - x_min), x_min, x_max, x

    @staticmethod
    def _mean_positive_distance(x):
        """Per input dimension: mean of |x_i - x_j| over ordered pairs with a positive distance
        (what lcgp.py:304-309 computes from d dense N x N matrices), in O(N log N) from the sort:
        sum_{i<j} (v_(j) - v_(i)) = sum_i v_(i) (2 i - N + 1)."""
        N = x.shape[0]
        out = torch.zeros(x.shape[1], dtype=DT)
        for j in range(x.shape[1]):
            v, _ = torch.sort(x[:, j])
            total = 2.0 * (v * (2.0 * torch.arange(N, dtype=DT) - (N - 1))).sum()
            _, counts = torch.unique_consecutive(v, return_counts=True)
            pairs = float(N) * N - float((counts.to(DT) ** 2).sum())
            out[j] = total / pairs if pairs > 0 else float('nan')
        return out

    @staticmethod
    def init_standard_x(x):
        """lcgp.py:295-310: (xs, x_min, x_max, x, xnorm)."""
        xs, x_min, x_max, x = LCGP._standardize_x(x)
        return xs, x_min, x_max, x, LCGP._mean_positive_distance(x)

    @property
    def xnorm(self):
        # unused by the model (SURVEY B-7); computed on first access
        if self._xnorm is None:
            self._xnorm = self._mean_positive_distance(self.x_orig)
        return self._xnorm

    def _prep_device(self):
        return torch.device('cuda', torch.cuda.current_device()) if self._device is None else torch.device(self._device)

    def _row_select_dev(self, Yd, center_d):
        """Nearest-rank median of the rows of Yd (or of |Yd - center|) by lcgp_prep_row_select."""
        p, m = int(Yd.shape[0]), int(Yd.shape[1])
        k = int(np.round((m - 1) * 0.5))
        out = torch.empty(p, dtype=DT, device=Yd.device)
        with torch.cuda.device(Yd.device):
            rc = _cabi.lib().lcgp_prep_row_select(Yd.data_ptr(), None if center_d is None else center_d.data_ptr(),
                                                  p, m, k, out.data_ptr(), _cabi.stream_ptr())
        _cabi.check(rc, 'lcgp_prep_row_select')
        return out

    def _standardize_dev(self, Yd, c_d, s_d, r_d=None, want=('Ys',)):
        """(Ys, YR, w) of lcgp_prep_standardize for the device matrix Yd; entries not in `want` are None."""
        p, n = int(Yd.shape[0]), int(Yd.shape[1])
        mk = lambda name, shape: torch.empty(shape, dtype=DT, device=Yd.device) if name in want else None
        Ys, YR, w = mk('Ys', (p, n)), mk('YR', (p, n)), mk('w', (p,))
        ptr = lambda t: None if t is None else t.data_ptr()
        with torch.cuda.device(Yd.device):
            rc = _cabi.lib().lcgp_prep_standardize(Yd.data_ptr(), c_d.data_ptr(), s_d.data_ptr(), ptr(r_d), p, n,
                                                   ptr(Ys), ptr(YR), ptr(w), _cabi.stream_ptr())
        _cabi.check(rc, 'lcgp_prep_standardize')
        return Ys, YR, w

    def _center_spread(self, Y, guard):
        if self.robust_mean and self._dev_prep:
            Yd = Y if Y.is_cuda else Y.to(self._prep_device(), DT).contiguous()
            c_d = self._row_select_dev(Yd, None)
            s_d = self._row_select_dev(Yd, c_d)
            c, s = c_d.cpu()[:, None], s_d.cpu()[:, None]
        elif self.robust_mean:
            c = _nearest_rank_median(Y)
            s = _nearest_rank_median(torch.abs(Y - c))
        else:
            if Y.is_cuda:
                Y = Y.cpu()
            c = Y.mean(dim=1, keepdim=True)
            s = Y.std(dim=1, keepdim=True, unbiased=False)
        if guard:   # lcgp.py:394 (rep path only)
            s = torch.where(s > 0, s, torch.ones_like(s))
        return c, s

    def init_standard_y(self, y):
        """lcgp.py:312-324."""
        if self._dev_prep:
            yd = y.to(self._prep_device(), DT).contiguous()
            c, s = self._center_spread(yd, guard=False)
            ys = self._standardize_dev(yd, c[:, 0].to(yd.device), s[:, 0].to(yd.device))[0]
            return ys.cpu(), c, s, y
        c, s = self._center_spread(y, guard=False)
        return (y - c) / s, c, s, y

    def _compute_center_spread_tf(self, Y):
        """lcgp.py:383-395 (name kept for drop-in use)."""
        return self._center_spread(Y if isinstance(Y, torch.Tensor) else torch.as_tensor(Y, dtype=DT), guard=True)

    # ------------------------------------------------------------------ replication
    def _get_raw_xy(self, x_raw=None, y_raw=None):
        xr = self.x_orig if x_raw is None else x_raw
        yr = self.y_orig if y_raw is None else y_raw
        xr = xr.numpy() if isinstance(xr, torch.Tensor) else np.asarray(xr, dtype=np.float64)
        yr = yr.numpy() if isinstance(yr, torch.Tensor) else np.asarray(yr, dtype=np.float64)
        assert xr.ndim == 2, 'x_raw must be (N, d)'
        assert yr.ndim == 2, 'y_raw must be (p, N)'
        assert yr.shape[1] == xr.shape[0], 'y_raw columns must match x_raw rows'
        return xr, yr, xr.shape[0], xr.shape[1], yr.shape[0]

    @staticmethod
    def _group_unique_rows_np(xr):
        """lcgp.py:349-356: lexicographically sorted unique rows, inverse map, counts."""
        xu, inv, cnt = np.unique(xr, axis=0, return_inverse=True, return_counts=True)
        return xu, np.asarray(inv).reshape(-1), cnt

    @staticmethod
    def _segments(inverse, n):
        """Stable argsort of the group ids and the n+1 segment boundaries (columns of group i, in their
        original order: order[offsets[i]:offsets[i+1]])."""
        order = np.argsort(inverse, kind='stable')
        offsets = np.concatenate([[0], np.cumsum(np.bincount(inverse, minlength=n))])
        return order, offsets

    @staticmethod
    def _compute_ybar_np(yr, inverse, n):
        """Replicate means on the raw scale (lcgp.py:358-367): the t-th replicate of every group is added in
        one vectorised step, t = 0, 1, ..., so each group is summed left to right in its original column
        order (what the reference's per-group numpy mean does for groups of fewer than 8 replicates)."""
        order, offsets = LCGP._segments(inverse, n)
        cnt = np.diff(offsets)
        sums = np.zeros((yr.shape[0], n), dtype=np.float64)
        for t in range(int(cnt.max())):
            g = np.nonzero(cnt > t)[0]
            sums[:, g] += yr[:, order[offsets[g] + t]]
        return sums / cnt[None, :]

    def _compute_ybar_dev(self, yr, inverse, n):
        """Same on the device (lcgp_prep_segment_mean); returns the device tensor."""
        order, offsets = self._segments(inverse, n)
        dev = self._prep_device()
        yd = torch.as_tensor(yr, dtype=DT).to(dev).contiguous()
        od = torch.as_tensor(order.astype(np.int32)).to(dev)
        fd = torch.as_tensor(offsets.astype(np.int32)).to(dev)
        p, N = int(yd.shape[0]), int(yd.shape[1])
        ybar = torch.empty((p, n), dtype=DT, device=dev)
        with torch.cuda.device(dev):
            rc = _cabi.lib().lcgp_prep_segment_mean(yd.data_ptr(), od.data_ptr(), fd.data_ptr(), p, N, n, ybar.data_ptr(),
                                                    _cabi.stream_ptr())
        _cabi.check(rc, 'lcgp_prep_segment_mean')
        return ybar

    def preprocess(self, y_raw=None, x_raw=None):
        """lcgp.py:397-426: 12-tuple of replication structures."""
        xr, yr, N, d, p = self._get_raw_xy(x_raw=x_raw, y_raw=y_raw)
        xu, inv, cnt = self._group_unique_rows_np(xr)
        n_unique = int(xu.shape[0])
        ybar_d = None
        if self._dev_prep:
            ybar_d = self._compute_ybar_dev(yr, inv, n_unique)
            ybar = ybar_d.cpu()
        else:
            ybar = torch.as_tensor(self._compute_ybar_np(yr, inv, n_unique), dtype=DT)
        x_unique = torch.as_tensor(xu, dtype=DT)
        x_unique_s = (x_unique - self.x_min) / (self.x_max - self.x_min)
        group_ids = torch.as_tensor(inv, dtype=torch.int32)
        r = torch.as_tensor(cnt.astype(np.int32))
        R = _LazyDiag(r)
        if ybar_d is not None:
            ybar_mean, ybar_std = self._compute_center_spread_tf(ybar_d)
            ybar_s = self._standardize_dev(ybar_d, ybar_mean[:, 0].to(ybar_d.device), ybar_std[:, 0].to(ybar_d.device))[0].cpu()
        else:
            ybar_mean, ybar_std = self._compute_center_spread_tf(ybar)
            ybar_s = (ybar - ybar_mean) / ybar_std
        return (x_unique, x_unique_s, group_ids, r, R, ybar, ybar_s, ybar_mean, ybar_std,
                _itensor(n_unique), _itensor(d), _itensor(p))

    @property
    def R(self):
        return torch.diag(self.r.to(DT))

    def _ensure_replication(self):
        if not self._rep_initialized:
            self.preprocess()
            self._rep_initialized = True

    # ------------------------------------------------------------------ basis
    def _get_phi_input(self):
        """lcgp.py:439-452."""
        if self.submethod != 'rep':
            return self.y
        if getattr(self, 'rep_standardize_ybar', True) and hasattr(self, 'ybar_s'):
            return self.ybar_s
        if hasattr(self, 'ybar'):
            return self.ybar
        return self.y

    def init_phi(self, var_threshold: float = None):
        """lcgp.py:454-485: phi = U_q sqrt(n) / s_q from the SVD of the (standardised) outputs."""
        Y = self._get_phi_input()
        n, p = int(self.n), int(self.p)
        U = s = None
        if self._dev_prep and self.q is not None and p >= 256 and 4 * int(self.q) <= p <= n:
            U, s = self._gram_basis(Y, int(self.q))        # None when the wanted components are not well enough separated
        if U is None:
            U, s = self._svd_basis(Y)
        if (self.q is None) and (var_threshold is None):
            q = p
        elif (self.q is None) and (var_threshold is not None):
            cum = np.cumsum(s.numpy() ** 2) / np.sum(s.numpy() ** 2)
            q = int(np.argmax(cum > var_threshold) + 1) if np.any(cum > var_threshold) else p
        else:
            q = int(self.q)
        assert U.shape[1] == min(n, p)
        phi = U[:, :q] * math.sqrt(n) / s[:q]
        diag_D = (phi ** 2).sum(dim=0)
        g = phi.T @ Y
        if self.verbose:   # the reference prints unconditionally (lcgp.py:482-483)
            print('======= VARIANCE OF G ======')
            print(g.var(dim=1, unbiased=False))
        return g, phi, diag_D, q

    def _gram_basis(self, Y, q):
        """Left singular vectors / singular values of Y (p x n, p <= n) from the p x p Gram matrix: Y Y^T = U S^2 U^T.
        The product (2 p^2 n flop, 64 GFLOP at config 4) runs on the device (a plain library DGEMM), the symmetric
        eigenproblem of size p on the host (LAPACK; rank 0 alone + broadcast when sharded): 0.3 s instead of the 2-4 s
        of the p x n SVD that dominated construction.  Squaring costs accuracy in the SMALL singular values
        (relative error ~ eps (s_1 / s_k)^2), so this is used only when q leading components are wanted and
        s_q >= s_1 / 300 (else None: the caller falls back to the SVD).  Signs are arbitrary, as with the SVD
        (SURVEY B-15: objective, gradient and predictions do not depend on them)."""
        dev = self._prep_device()
        p = int(Y.shape[0])
        if self._world == 1 or self._rank == 0:
            Yd = Y.to(dev, DT)
            G = (Yd @ Yd.T).cpu()
            G = 0.5 * (G + G.T)
            prev = torch.get_num_threads()
            if self._world > 1:
                torch.set_num_threads(max(prev, (os.cpu_count() or 1) - self._world + 1))
            try:
                lam, V = torch.linalg.eigh(G)
            finally:
                torch.set_num_threads(prev)
            lam, V = torch.flip(lam, dims=[0]), torch.flip(V, dims=[1])
            s = torch.sqrt(torch.clamp(lam, min=0.0))
            ok = bool(s[q - 1] * 300.0 >= s[0]) and bool(torch.isfinite(s).all())
            buf = torch.cat([torch.tensor([1.0 if ok else 0.0], dtype=DT), V.reshape(-1), s])
        if self._world > 1:
            cdev = self._collective_device()
            buf = buf.to(cdev) if self._rank == 0 else torch.empty(1 + p * p + p, dtype=DT, device=cdev)
            torch.distributed.broadcast(buf, src=0)
            buf = buf.cpu()
        if float(buf[0]) == 0.0:
            return None, None
        return buf[1:1 + p * p].reshape(p, p), buf[1 + p * p:]

    def _svd_basis(self, Y):
        """Left singular vectors and singular values of Y (p x n) by LAPACK on the host.  With the latents
        sharded over ranks, rank 0 alone factors (with the host threads the other ranks would otherwise
        fight over) and broadcasts U, s: N concurrent SVDs on one host were the dominant construction
        cost, and every rank then holds bit-identical phi."""
        if self._world == 1:
            U, s, _ = torch.linalg.svd(Y, full_matrices=False)
            return U, s
        import os
        dist = torch.distributed
        dev = self._collective_device()
        r = min(Y.shape)
        if self._rank == 0:
            prev = torch.get_num_threads()
            torch.set_num_threads(max(prev, (os.cpu_count() or 1) - self._world + 1))
            try:
                U, s, _ = torch.linalg.svd(Y, full_matrices=False)
            finally:
                torch.set_num_threads(prev)
            buf = torch.cat([U.reshape(-1), s]).to(dev)
        else:
            buf = torch.empty(Y.shape[0] * r + r, dtype=DT, device=dev)
        dist.broadcast(buf, src=0)
        buf = buf.cpu()
        return buf[:Y.shape[0] * r].reshape(Y.shape[0], r), buf[Y.shape[0] * r:]

    # ------------------------------------------------------------------ parameters
    def init_params(self):
        """lcgp.py:490-513."""
        x = self.x.numpy()
        d = int(self.d)
        llmb = np.exp(0.5 * np.log(d) + np.log(np.std(x, axis=0)))
        ynp = self.y.numpy()
        lsig = np.zeros(len(self.diag_error_structure))
        col = 0
        for k, width in enumerate(self.diag_error_structure):
            lsig[k] = np.log(np.var(ynp[col:col + width]))
            col += width
        self.lLmb.assign(np.tile(llmb, self.q).reshape(self.q, d))
        self.lLmb0.assign(np.ones(self.q))
        self.lnugGPs.assign(np.exp(-10.0) * np.ones(self.q))
        self.lsigma2s.assign(lsig)

    @property
    def trainable_variables(self):
        # attribute-name order of the reference's tf.Module: lLmb, lLmb0, lnugGPs, lsigma2s
        return [self.lLmb.unconstrained, self.lLmb0.unconstrained, self.lnugGPs.unconstrained,
                self.lsigma2s.unconstrained]

    def _get_param_graph(self):
        idx = getattr(self, '_err_index', None)
        if idx is None or idx.numel() != int(self.p):
            # output j -> its error group (get_param's expansion, lcgp.py:521-530), built once:
            # torch.repeat_interleave with a repeats tensor costs ~1 ms per call
            reps = torch.as_tensor(self.diag_error_structure, dtype=torch.long)
            idx = self._err_index = torch.repeat_interleave(torch.arange(len(self.diag_error_structure)), reps)
        return (self.lLmb.value(), self.lLmb0.value(), self.lsigma2s.value()[idx], self.lnugGPs.value())

    def get_param(self):
        """lcgp.py:515-532: (lLmb, lLmb0, lsigma2s expanded to a p-vector, lnugGPs), detached."""
        return tuple(t.detach() for t in self._get_param_graph())

    # ------------------------------------------------------------------ engine
    def _problem_data(self):
        """Constant arrays of the objective in the form the C-ABI takes (include/lcgp_b200.h)."""
        if self.submethod == 'rep':
            r = self.r.to(DT)
            X = self.x_unique_s
            Ybar = self.ybar_s if self.rep_standardize_ybar else self.ybar
            t = self.ybar_std[:, 0] if self.rep_standardize_ybar else torch.ones(int(self.p), dtype=DT)
            scale = 1.0 / float(self.n)
        else:
            r = torch.ones(int(self.n), dtype=DT)
            X = self.x
            Ybar = self.y
            t = torch.ones(int(self.p), dtype=DT)
            scale = 1.0
        if self._dev_prep:    # YR and w by lcgp_prep_standardize with centre 0 / spread 1 (exact), on the device
            dev = self._prep_device()
            p = int(self.p)
            zero, one = torch.zeros(p, dtype=DT, device=dev), torch.ones(p, dtype=DT, device=dev)
            _, YR, w = self._standardize_dev(Ybar.to(dev, DT).contiguous(), zero, one, r.to(dev), want=('YR', 'w'))
        else:
            YR = Ybar * r[None, :]
            w = (YR * Ybar).sum(dim=1)
        return dict(n=int(self.n), d=int(self.d), p=int(self.p), X=X, sr=torch.sqrt(r), YR=YR,
                    w=w, t=t, scale=scale, sum_log_r=float(torch.log(r).sum()))

    @property
    def engine(self):
        if self._engine is None:
            data = self._problem_data()
            idx = self._local_idx
            if idx.numel() == 0:
                self._engine = False      # this rank owns no latent
            else:
                factory = self._engine_factory or CudaEngine
                kw = {} if self._engine_factory else dict(device=self._device, stream_groups=self._stream_groups)
                self._engine = factory(phi_loc=self.phi[:, idx], D_loc=self.diag_D[idx],
                                       include_host_terms=(self._rank == 0), **data, **kw)
        return self._engine

    def _evaluate_sharded(self, lLmb, lLmb0, lsig_p, lnug, need_grad, events=None):
        """Multi-rank evaluation: local latents on this rank's engine, then ONE all-reduce of the flat vector
        [objective | d/d lsigma2 (p) | d/d lLmb (q x d) | d/d lLmb0 (q) | d/d lnugGPs (q) | failed latents]
        (each rank fills only its own latents' rows; lcgp_pack_sharded writes it in one launch into a persistent
        buffer).  Returns the reduced vector on the collective device without synchronising the host.  The last
        slot makes a Cholesky failure collective: every rank sees the same count and raises together."""
        q, d, p = int(self.q), int(self.d), int(self.p)
        idx = self._local_idx
        eng = self.engine
        ql = idx.numel()
        nflat = 1 + p + q * d + 2 * q + 1
        if eng is False:
            flat = self._flat_buffer(nflat, self._collective_device())
            flat.zero_()
        elif hasattr(eng, 'pack_sharded'):
            eng.evaluate_device(lLmb[idx], lLmb0[idx], lnug[idx], lsig_p, need_grad, events)
            flat = self._flat_buffer(nflat, eng.device)
            if getattr(self, '_loc_of', None) is None:
                loc = torch.full((q,), -1, dtype=torch.int32)
                loc[idx] = torch.arange(ql, dtype=torch.int32)
                self._loc_of = loc.to(eng.device)
            eng.pack_sharded(self._loc_of, q, flat)
        else:       # engines without the packing kernel (the CPU stand-in of the gloo tests)
            out = eng.evaluate(lLmb[idx], lLmb0[idx], lnug[idx], lsig_p, need_grad, events)
            flat = torch.zeros(nflat, dtype=DT, device=out.device)
            flat[:1 + p] = out[:1 + p]
            if need_grad:
                o = 1 + p
                di = idx.to(out.device)
                flat[o:o + q * d].view(q, d)[di] = out[o:o + ql * d].view(ql, d)
                flat[o + q * d:o + q * d + q][di] = out[o + ql * d:o + ql * d + ql]
                flat[o + q * d + q:o + q * d + 2 * q][di] = out[o + ql * d + ql:o + ql * d + 2 * ql]
        torch.distributed.all_reduce(flat)
        return flat

    def _flat_buffer(self, nflat, device):
        buf = getattr(self, '_flat_buf', None)
        if buf is None or buf.numel() != nflat or buf.device != torch.device(device):
            buf = self._flat_buf = torch.zeros(nflat, dtype=DT, device=device)
        return buf

    def _evaluate(self, lLmb, lLmb0, lsig_p, lnug, need_grad, events=None):
        """All-latent objective (+ gradient wrt the constrained values) as CPU tensors."""
        q, d, p = int(self.q), int(self.d), int(self.p)
        eng = self.engine
        self.n_evals += 1
        if self._world == 1:
            flat = eng.evaluate(lLmb, lLmb0, lnug, lsig_p, need_grad, events)[:1 + p + q * d + 2 * q]
        else:
            flat = self._evaluate_sharded(lLmb, lLmb0, lsig_p, lnug, need_grad, events).cpu()
            if float(flat[-1]) != 0.0:     # identical on every rank after the all-reduce: all ranks raise together
                raise RuntimeError(f'lcgp_b200: Cholesky failed for {int(flat[-1])} latent(s) on some rank '
                                   f'(A_k = I + d_k R^1/2 C_k R^1/2 has eigenvalues >= 1, so this means NaN/inf inputs)')
        self._factor_key = self._key(lLmb, lLmb0, lsig_p, lnug)
        val = flat[0].clone()
        if not need_grad:
            return val, None
        o = 1 + p
        return val, (flat[o:o + q * d].reshape(q, d).clone(), flat[o + q * d:o + q * d + q].clone(),
                     flat[1:1 + p].clone(), flat[o + q * d + q:o + q * d + 2 * q].clone())

    def _collective_device(self):
        be = torch.distributed.get_backend()
        return torch.device('cuda', torch.cuda.current_device()) if be == 'nccl' else torch.device('cpu')

    @staticmethod
    def _key(*ts):
        return b''.join(t.detach().cpu().numpy().tobytes() for t in ts)

    # ------------------------------------------------------------------ losses
    def loss(self):
        """lcgp.py:542-549."""
        try:
            return self.submethod_loss_map[self.submethod]()
        except KeyError:
            raise ValueError("Invalid submethod. Choices are 'full' or 'rep'.")

    def _neglpost(self):
        lLmb, lLmb0, lsig_p, lnug = self._get_param_graph()
        return _NegLogPost.apply(self, lLmb, lLmb0, lsig_p, lnug)

    def neglpost_rep(self):
        """Replicated negative log posterior, divided by n (lcgp.py:554-630)."""
        return self._neglpost()

    def neglpost(self):
        """Negative log posterior of the unreplicated model (lcgp.py:635-666)."""
        return self._neglpost()

    def loss_and_grad(self):
        """(objective, flat gradient wrt the unconstrained variables in trainable_variables order)."""
        tv = self.trainable_variables
        for v in tv:
            v.grad = None
        val = self.loss()
        val.backward()
        return float(val.detach()), torch.cat([v.grad.reshape(-1) for v in tv]).numpy().copy()

    def grad_phi(self):
        """Gradient of loss() with respect to the latent basis phi (p x q), with diag_D following phi
        (d_k = sum_j phi_jk^2).  The reference keeps phi constant (lcgp.py:164); this serves callers that train the
        basis.  Sharded like the objective: each rank computes its latents' columns, one all-reduce gathers them."""
        lLmb, lLmb0, lsig_p, lnug = self.get_param()
        self._evaluate(lLmb, lLmb0, lsig_p, lnug, need_grad=True)
        q, p = int(self.q), int(self.p)
        eng = self.engine
        loc = eng.grad_phi(lsig_p) if eng is not False else None
        if self._world == 1:
            return loc.cpu()
        dev = self._collective_device()
        full = torch.zeros((p, q), dtype=DT, device=dev)
        if loc is not None:
            full[:, self._local_idx.to(dev)] = loc.to(dev)
        torch.distributed.all_reduce(full)
        return full.cpu()

    def _flat_get(self):
        return torch.cat([v.detach().reshape(-1) for v in self.trainable_variables]).numpy().copy()

    def _flat_set(self, vec):
        vec = torch.as_tensor(np.asarray(vec, dtype=np.float64), dtype=DT)
        o = 0
        with torch.no_grad():
            for v in self.trainable_variables:
                m = v.numel()
                v.copy_(vec[o:o + m].reshape(v.shape))
                o += m

    # ------------------------------------------------------------------ fit
    def fit(self, verbose=False, optimizer: str = 'L-BFGS-B', **options):
        """lcgp.py:537-540.  optimizer='L-BFGS-B' is SciPy's, as driven by gpflow.optimizers.Scipy in
        the reference; optimizer='torch-lbfgs' is torch.optim.LBFGS with strong-Wolfe line search.
        Both run on the host and call the CUDA objective once per closure evaluation."""
        self._invalidate_aux()
        if optimizer == 'L-BFGS-B':
            import scipy.optimize

            def fun(v):
                self._flat_set(v)
                f, g = self.loss_and_grad()
                if verbose:
                    print(f'  eval {self.n_evals}: loss {f:.10g}  |g| {np.linalg.norm(g):.3e}')
                return f, g
            res = scipy.optimize.minimize(fun, self._flat_get(), jac=True, method='L-BFGS-B',
                                          options=options or None)
            self._flat_set(res.x)
            self.opt_result = res
        elif optimizer == 'torch-lbfgs':
            kw = dict(lr=1.0, max_iter=500, history_size=10, line_search_fn='strong_wolfe',
                      tolerance_grad=1e-9, tolerance_change=1e-12)
            kw.update(options)
            opt = torch.optim.LBFGS(self.trainable_variables, **kw)

            def closure():
                opt.zero_grad()
                val = self.loss()
                val.backward()
                return val
            opt.step(closure)
            self.opt_result = opt.state_dict()['state']
        else:
            raise ValueError("optimizer must be 'L-BFGS-B' or 'torch-lbfgs'")
        self._invalidate_aux()
        return

    # ------------------------------------------------------------------ state save / restore
    def state_dict(self):
        """Fitted state (the reference has no checkpointing): the four parameter tensors
        in unconstrained form plus the data fingerprint they belong to."""
        return {'format': 'lcgp_b200/1', 'submethod': self.submethod, 'q': int(self.q), 'n': int(self.n),
                'd': int(self.d), 'p': int(self.p), 'diag_error_structure': list(self.diag_error_structure),
                'diag_D': self.diag_D.clone(), 'x_checksum': self._x_checksum(),
                'params': {k: getattr(self, k).unconstrained.detach().clone()
                           for k in ('lLmb', 'lLmb0', 'lsigma2s', 'lnugGPs')}}

    def _x_checksum(self):
        X = self.x_unique_s if self.submethod == 'rep' else self.x
        wgt = torch.arange(1, X.numel() + 1, dtype=DT).reshape(X.shape)
        return float((X * wgt).sum())

    def load_state_dict(self, state):
        """Restore parameters saved by state_dict() into a model built on the same data."""
        for k in ('submethod', 'q', 'n', 'd', 'p'):
            if state[k] != (getattr(self, k) if k == 'submethod' else int(getattr(self, k))):
                raise ValueError(f'state_dict mismatch in {k!r}: {state[k]} vs {getattr(self, k)}')
        if list(state['diag_error_structure']) != list(self.diag_error_structure) or \
                not torch.allclose(state['diag_D'], self.diag_D, rtol=1e-10, atol=0) or \
                abs(state['x_checksum'] - self._x_checksum()) > 1e-9 * max(1.0, abs(state['x_checksum'])):
            raise ValueError('state_dict was saved for different data / error structure')
        with torch.no_grad():
            for k, v in state['params'].items():
                getattr(self, k).unconstrained.copy_(v)
        self._invalidate_aux()
        return self

    # ------------------------------------------------------------------ aux predictive quantities
    def _invalidate_aux(self):
        self._factor_key = None
        self.CinvMs = torch.full((int(self.q), int(self.n)), float('nan'), dtype=DT)
        self.mks = torch.full((int(self.q), int(self.n)), float('nan'), dtype=DT)
        self.Tks = None
        self.Ths = None

    def _refresh_factor(self):
        """Make sure the workspace holds the factor for the current parameters."""
        lLmb, lLmb0, lsig_p, lnug = self.get_param()
        if self._factor_key != self._key(lLmb, lLmb0, lsig_p, lnug):
            self._evaluate(lLmb, lLmb0, lsig_p, lnug, need_grad=False)
        return lLmb, lLmb0, lsig_p, lnug

    def _gather_rows(self, local, width):
        """(q_loc x width) per rank -> (q x width) on the CPU, rows placed by latent index."""
        q = int(self.q)
        if self._world == 1:
            return local.cpu()
        dev = self._collective_device()
        full = torch.zeros((q, width), dtype=DT, device=dev)
        if local is not None:
            full[self._local_idx.to(dev)] = local.to(dev)
        torch.distributed.all_reduce(full)
        return full.cpu()

    def compute_aux_predictive_quantities(self):
        """lcgp.py:685-726 / 728-803: CinvMs (= alpha_k), mks (rep) and lazy Tks / Ths."""
        self._refresh_factor()
        eng = self.engine
        a, m = eng.aux() if eng is not False else (None, None)
        n = int(self.n)
        self.CinvMs = self._gather_rows(a, n)
        self.mks = self._gather_rows(m, n)
        lsig_p = self.get_param()[2]
        sinv = torch.exp(-0.5 * lsig_p)
        if self.submethod == 'rep':
            if self.rep_standardize_ybar:
                sinv = sinv * self.ybar_std[:, 0]
            q, p = int(self.q), int(self.p)
            # lcgp.py:754 as coded broadcasts only for q == p (SURVEY B-3); otherwise divide columns
            self.psi_c = self.phi.T / sinv[:, None] if (q == p or p == 1) else self.phi.T / sinv[None, :]
            self.Tks = _LazyOperators(self, 'Tks')
            self.Ths = None
        else:
            self.Ths = _LazyOperators(self, 'Ths')

    def _compute_aux_predictive_quantities_rep(self):
        self.compute_aux_predictive_quantities()

    def _aux_is_stale(self):
        return bool(torch.isnan(self.CinvMs).any()) or (self.Tks is None and self.Ths is None)

    # ------------------------------------------------------------------ prediction
    def predict(self, x0, return_fullcov=False):
        """lcgp.py:671-680."""
        x0 = _as_tensor2d(x0)
        try:
            call = self.submethod_predict_map[self.submethod]
        except KeyError as e:
            print(e)
            raise KeyError('Invalid submethod.  Choices are \'full\' or \'rep\'.')
        res = call(x0=x0, return_fullcov=return_fullcov)
        return tuple(t.detach() if t is not None else None for t in res)

    def _gather_latent_moments(self, gh, gv, width):
        """This rank's (q_loc x width) latent means / variances -> the full (q x width) pair ON THE DEVICE of the
        collective: ONE all_gather of the stacked pair (rows padded to the largest q_loc) instead of two all-reduces."""
        q = int(self.q)
        if self._world == 1:
            return gh, gv
        W = self._world
        dev = self._collective_device()
        qmax = -(-q // W)
        loc = torch.zeros((2, qmax, width), dtype=DT, device=dev)
        if gh is not None:
            loc[0, :gh.shape[0]] = gh.to(dev)
            loc[1, :gv.shape[0]] = gv.to(dev)
        allb = torch.empty((W * 2, qmax, width), dtype=DT, device=dev)      # concatenation along dim 0 (gloo wants that shape)
        torch.distributed.all_gather_into_tensor(allb, loc)
        allb = allb.view(W, 2, qmax, width)
        k = torch.arange(q, device=dev)
        full = allb[k % W, :, k // W]                      # latent k lives on rank k % W at local row k // W
        return full[:, 0].contiguous(), full[:, 1].contiguous()

    def _predict_latents(self, x0, Xtrain, chunk=2048, keep_on_device=False):
        if self._aux_is_stale():
            self.compute_aux_predictive_quantities()
        lLmb, lLmb0, lsig_p, lnug = self._refresh_factor()
        x0s = (x0 - self.x_min) / (self.x_max - self.x_min)                 # lcgp.py:822 / :877
        same = x0s.shape == Xtrain.shape and bool(torch.equal(x0s, Xtrain))  # covmat.py:46-49
        idx = self._local_idx
        eng = self.engine
        n0 = int(x0s.shape[0])
        if same:
            chunk = n0    # the nugget-on-the-diagonal quirk needs global row indices
        gh, gv = [], []
        for s in range(0, n0, chunk):
            if eng is not False:
                a, b = eng.predict_latents(lLmb[idx], lLmb0[idx], lnug[idx], x0s[s:s + chunk].contiguous(), same)
            else:
                a = b = None
            a, b = self._gather_latent_moments(a, b, min(chunk, n0 - s))
            gh.append(a)
            gv.append(b)
        ghat_d = gh[0] if len(gh) == 1 else torch.cat(gh, dim=1)
        gvar_d = gv[0] if len(gv) == 1 else torch.cat(gv, dim=1)
        self._ghat_dev, self._gvar_dev = ghat_d, gvar_d
        self._ghat = self._gvar = None                    # CPU copies on demand (properties ghat / gvar)
        if keep_on_device and ghat_d.is_cuda:
            return ghat_d, gvar_d, lsig_p
        return self.ghat, self.gvar, lsig_p

    def _latent_moments_host(self):
        """ghat / gvar (q x n0) of the last prediction as CPU tensors (reference attributes, lcgp.py:899-900)."""
        if self._ghat is None and self._ghat_dev is not None:
            self._ghat, self._gvar = self._ghat_dev.cpu(), self._gvar_dev.cpu()
        return self._ghat, self._gvar

    @property
    def ghat(self):
        return self._latent_moments_host()[0]

    @ghat.setter
    def ghat(self, v):
        self._ghat, self._ghat_dev = v, None

    @property
    def gvar(self):
        return self._latent_moments_host()[1]

    @gvar.setter
    def gvar(self, v):
        self._gvar, self._gvar_dev = v, None

    def _output_maps(self, Psi, ghat, gvar, noise_var, scale, shift):
        """lcgp.py:915-926 / :840-848: (ypred, ypredvar, yconfvar), each p x n0, as CPU tensors.  With the latent moments
        on a CUDA device the three products and the un-standardisation run there (lcgp_predict_outputs) and only the
        results come back (pinned host memory); otherwise (test engines on the CPU) in torch on the host."""
        p, q, n0 = int(self.p), int(self.q), int(ghat.shape[1])
        if ghat.is_cuda and _cabi.available():
            dev = ghat.device
            c = lambda t: None if t is None else t.to(dev, DT).contiguous()
            Psi_d, nv_d, sc_d, sh_d = c(Psi), c(noise_var), c(scale), c(shift)
            outs = [torch.empty((p, n0), dtype=DT, device=dev) for _ in range(3)]
            ptr = lambda t: None if t is None else t.data_ptr()
            with torch.cuda.device(dev):
                rc = _cabi.lib().lcgp_predict_outputs(Psi_d.data_ptr(), ghat.data_ptr(), gvar.data_ptr(), nv_d.data_ptr(),
                                                      ptr(sc_d), ptr(sh_d), p, q, n0, outs[0].data_ptr(), outs[1].data_ptr(),
                                                      outs[2].data_ptr(), _cabi.stream_ptr())
                _cabi.check(rc, 'lcgp_predict_outputs')
                host = [torch.empty((p, n0), dtype=DT, pin_memory=True) for _ in range(3)]
                for h, o in zip(host, outs):
                    h.copy_(o, non_blocking=True)
                torch.cuda.current_stream().synchronize()
            return tuple(host)
        predmean = Psi @ ghat
        confvar = (Psi ** 2) @ gvar
        predvar = confvar + noise_var[:, None]
        sc = torch.ones(p, dtype=DT) if scale is None else scale
        sh = torch.zeros(p, dtype=DT) if shift is None else shift
        return predmean * sc[:, None] + sh[:, None], predvar * sc[:, None] ** 2, confvar * sc[:, None] ** 2

    def predict_full(self, x0, return_fullcov=False):
        """lcgp.py:808-859."""
        ghat, gvar, lsig_p = self._predict_latents(x0, self.x, keep_on_device=True)
        sig2 = torch.exp(lsig_p)
        Psi = self.phi * torch.sqrt(sig2)[:, None]                          # (p, q) = psi^T of lcgp.py:838
        ypred, ypredvar, yconfvar = self._output_maps(Psi, ghat, gvar, sig2, self.ystd[:, 0], self.ymean[:, 0])
        if return_fullcov:
            # lcgp.py:850-857: n0 rank-q updates of a diagonal, written by one kernel (csrc/fullcov.cu)
            full = _fullcov_cuda(Psi.T.contiguous(), self._latent_moments_host()[1], sig2, self.ystd[:, 0], self._device)
            return ypred, ypredvar, yconfvar, full
        return ypred, ypredvar, yconfvar

    def predict_rep(self, x0, return_fullcov=False):
        """lcgp.py:864-930."""
        ghat, gvar, lsig_p = self._predict_latents(x0, self.x_unique_s, keep_on_device=True)
        sigma_var = torch.exp(lsig_p)
        sigma_sqrt = torch.sqrt(sigma_var)
        scale = shift = None
        if self.rep_standardize_ybar:
            std = self.ybar_std[:, 0]
            sigma_sqrt = sigma_sqrt / std
            sigma_var = sigma_var / std ** 2
            scale, shift = std, self.ybar_mean[:, 0]
        Psi = self.phi * sigma_sqrt[:, None]
        ypred, ypredvar, yconfvar = self._output_maps(Psi, ghat, gvar, sigma_var, scale, shift)
        if return_fullcov:
            return ypred, ypredvar, yconfvar, None
        return ypred, ypredvar, yconfvar


def _fullcov_cuda(psi, gvar, sig2, ystd, device=None, chunk_bytes=4 << 30):
    """(n0, p, p) predictive covariance on the CPU, computed on the GPU in chunks of test points."""
    _cabi.require_cuda()
    L = _cabi.lib()
    dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    q, p = int(psi.shape[0]), int(psi.shape[1])
    n0 = int(gvar.shape[1])
    c = lambda a: a.to(dev, DT).contiguous()
    psi_d, sig_d, sv_d = c(psi), c(sig2), c(ystd)
    step = max(1, min(n0, int(chunk_bytes // (8 * p * p)), 1 << 19))
    out = torch.empty((n0, p, p), dtype=DT)
    with torch.cuda.device(dev):
        for s in range(0, n0, step):
            w = min(step, n0 - s)
            gv = c(gvar[:, s:s + w])
            buf = torch.empty((w, p, p), dtype=DT, device=dev)
            rc = L.lcgp_predict_fullcov(psi_d.data_ptr(), gv.data_ptr(), sig_d.data_ptr(), sv_d.data_ptr(), q, p, w,
                                        buf.data_ptr(), _cabi.stream_ptr())
            _cabi.check(rc, 'lcgp_predict_fullcov')
            out[s:s + w] = buf.cpu()
    return out


class _LazyDiag:
    """diag(r) (lcgp.py:378) without the n x n allocation until someone asks for it."""

    def __init__(self, r):
        self._r = r
        self.shape = (r.shape[0], r.shape[0])

    def materialize(self):
        return torch.diag(self._r.to(DT))

    def numpy(self):
        return self.materialize().numpy()


class _LazyOperators:
    """Stand-in for the reference's q x n x n `Tks` / `Ths` tensors (lcgp.py:760, 700): 16 GB at
    n = 8000, q = 32.  Indexing with k rebuilds operator k from the factor in the workspace:
    Tks[k] = d_k (sqrt r sqrt r^T) o A_k^{-1}   (== C^-1 - C^-1 (C^-1 + d_k R)^-1 C^-1, lcgp.py:783-788)
    Ths[k] = sqrt(d_k) L_k^{-T}                 (Th Th^T == (C_k + I/d_k)^{-1} as in lcgp.py:709-715)"""

    def __init__(self, model, kind):
        self._m, self._kind = model, kind
        self.shape = (int(model.q), int(model.n), int(model.n))

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, k):
        m = self._m
        k = int(k)
        m._refresh_factor()
        if m._world == 1:
            Ainv = m.engine.ainv(k).cpu()
        else:       # sharded: the rank that owns latent k rebuilds A_k^-1 and broadcasts it (a collective: every rank must index)
            W, n = m._world, int(m.n)
            dev = m._collective_device()
            if k % W == m._rank:
                buf = m.engine.ainv(k // W).to(dev)
            else:
                buf = torch.empty((n, n), dtype=DT, device=dev)
            torch.distributed.broadcast(buf, src=k % W)
            Ainv = buf.cpu()
        dk = m.diag_D[k]
        if self._kind == 'Tks':
            sr = torch.sqrt(m.r.to(DT))
            return dk * (sr[:, None] * sr[None, :]) * Ainv
        # symmetric square root is not unique; return the Cholesky-type factor of d_k A^{-1}
        return torch.linalg.cholesky(dk * Ainv)

    def numpy(self):
        return torch.stack([self[k] for k in range(self.shape[0])]).numpy()
