"""Batched fitting of independent emulators / L-BFGS restarts (BASELINE.json config 5).

The emulators are independent, so there is no data-path collective ("replicas only"): emulator i is fitted
by rank `i mod world` on that rank's GPU and only the fitted parameters are gathered at the end.

Inside a rank the emulators advance in LOCK-STEP: every emulator runs its own SciPy L-BFGS-B state machine (the very
routine `scipy.optimize.minimize` -- i.e. the reference's `gpflow.optimizers.Scipy`, lcgp.py:537-540 -- drives, see
lbfgsb.py), and the objective + gradient evaluations all machines are waiting for are served by ONE device call over
all their latents (`lcgp_problem.n_emu`: E emulators x q latents, e.g. 8 x 8 matrices of n = 1024 per launch).  One
small emulator cannot fill 148 SMs and one Python thread per emulator saturates on the interpreter lock (round 1:
~1000 evaluations/s however many threads); a batch of E emulators costs the device little more than one (the
factorisation is a dependency chain, see csrc/potrf_pll.cu) and the host one vectorised chain rule per step.
Emulators whose shapes differ (or `engine='threads'`) fall back to one host thread + CUDA stream + graph per emulator.
"""
from __future__ import annotations

import os
import threading
import time
from typing import Callable, Sequence

import numpy as np
import torch

from . import _cabi
from .lbfgsb import LbfgsbMachine
from .model import DT, LCGP


def perturbed_restart(seed: int, lo: float = 0.5, hi: float = 2.0) -> Callable[[LCGP], None]:
    """Start-point perturbation for restarts: length-scales multiplied by U(lo, hi) factors."""
    def apply(model: LCGP):
        rng = np.random.default_rng(seed)
        model.lLmb.assign(model.lLmb.numpy() * rng.uniform(lo, hi, model.lLmb.numpy().shape))
    return apply


class BatchedEngine:
    """Constant data of E emulators of identical (n, d, p, q, submethod) on one device + the batched C-ABI call.
    `set_active` restricts the call to a subset of the emulators (those whose optimizer is still running): their
    constants are gathered into the leading slots of fixed device buffers, so the workspace, the pinned staging and
    (for small batches) the captured graphs stay valid."""

    KEYS = ('X', 'sr', 'YR', 'w', 't', 'phi', 'D', 'consts')

    def __init__(self, models: Sequence[LCGP], device=None):
        _cabi.require_cuda()
        self.lib = _cabi.lib()
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        m0 = models[0]
        self.E_all = E = len(models)
        self.n, self.d, self.p, self.q = int(m0.n), int(m0.d), int(m0.p), int(m0.q)
        if self.d > _cabi.MAX_D:
            raise ValueError(f'lcgp_b200 supports input dimension d <= {_cabi.MAX_D}')
        data = [m._problem_data() for m in models]
        dev = self.device
        st = lambda key: torch.stack([dd[key].to(dev, DT) for dd in data]).contiguous()
        self._all = dict(X=st('X'), sr=st('sr'), YR=st('YR'), w=st('w'), t=st('t'),
                         phi=torch.stack([m.phi.to(dev, DT) for m in models]).contiguous(),           # E x p x q
                         D=torch.stack([m.diag_D.to(dev, DT) for m in models]).contiguous(),          # E x q
                         consts=torch.tensor([[dd['scale'], dd['sum_log_r']] for dd in data], dtype=DT, device=dev))
        self._consts_host = [(float(dd['scale']), float(dd['sum_log_r'])) for dd in data]
        self._buf = {k: torch.empty_like(v) for k, v in self._all.items()}
        self.ws_bytes = int(self.lib.lcgp_workspace_bytes_batched(self.n, self.d, self.p, self.q, E))
        self.ws = torch.empty(self.ws_bytes // 8, dtype=DT, device=dev)
        self.out_len = int(self.lib.lcgp_out_len(self.p, self.d, self.q))
        q_all = E * self.q
        self.h_par = torch.empty(q_all * self.d + 2 * q_all + E * self.p, dtype=DT).pin_memory()
        self.h_out = torch.empty(E * self.out_len, dtype=DT).pin_memory()
        self.h_info = torch.zeros(q_all, dtype=torch.int32).pin_memory()
        self.small = _cabi.padded(self.n) // _cabi.NB <= 16
        self._plans = {}
        self.n_calls = 0
        self.set_active(list(range(E)))

    def set_active(self, slots):
        """Evaluate only emulators `slots` (indices into the constructor's list) from now on, in that order."""
        self.active = list(slots)
        a = self.E = len(self.active)
        idx = torch.as_tensor(self.active, dtype=torch.long, device=self.device)
        b = self._buf
        for k in self.KEYS:
            b[k][:a].copy_(self._all[k].index_select(0, idx))
        sc, slr = self._consts_host[self.active[0]] if a == 1 else (0.0, 0.0)
        self.prob = _cabi.Problem(n=self.n, d=self.d, p=self.p, q_loc=a * self.q, include_host_terms=1,
                                  n_emu=a if a > 1 else 0, scale=sc, sum_log_r=slr, X=b['X'].data_ptr(),
                                  sr=b['sr'].data_ptr(), YR=b['YR'].data_ptr(), w=b['w'].data_ptr(), t=b['t'].data_ptr(),
                                  phi=b['phi'].data_ptr(), D=b['D'].data_ptr(),
                                  emu_consts=b['consts'].data_ptr() if a > 1 else None)
        self.h2d_bytes = (a * self.q * (self.d + 2) + a * self.p) * 8
        self.d2h_bytes = a * self.out_len * 8 + a * self.q * 4

    def evaluate(self, lLmb, lLmb0, lnug, lsig_p, with_grad=True):
        """Constrained parameters of the ACTIVE emulators (E x q x d, E x q, E x q, E x p; CPU) -> their `out` blocks
        (E x out_len, a CPU view of the pinned result buffer: copy what must survive the next call)."""
        E, q, d = self.E, self.q, self.d
        ql = E * q
        npar = ql * d + 2 * ql + E * self.p
        h = self.h_par
        h[:ql * d].copy_(lLmb.reshape(-1))
        h[ql * d:ql * d + ql].copy_(lLmb0.reshape(-1))
        h[ql * d + ql:ql * d + 2 * ql].copy_(lnug.reshape(-1))
        h[ql * d + 2 * ql:npar].copy_(lsig_p.reshape(-1))
        flags = int(bool(with_grad))
        self.n_calls += 1
        import ctypes as C
        with torch.cuda.device(self.device):
            if self.small and ql < 64:      # launch-bound: replay a captured graph (one per active count)
                plan = self._plans.get((E, flags))
                if plan is None:
                    hdl = C.c_void_p()
                    _cabi.check(self.lib.lcgp_plan_create(self.prob, self.ws.data_ptr(), self.ws_bytes, h.data_ptr(),
                                                          self.h_out.data_ptr(), self.h_info.data_ptr(), flags, C.byref(hdl)),
                                'lcgp_plan_create')
                    plan = self._plans[(E, flags)] = hdl
                _cabi.check(self.lib.lcgp_plan_run(plan, _cabi.stream_ptr()), 'lcgp_plan_run')
            else:
                base = h.data_ptr()
                _cabi.check(self.lib.lcgp_nll_grad_host(self.prob, base, base + 8 * ql * d, base + 8 * (ql * d + ql),
                                                        base + 8 * (ql * d + 2 * ql), self.ws.data_ptr(), self.ws_bytes,
                                                        self.h_out.data_ptr(), self.h_info.data_ptr(), flags, None,
                                                        _cabi.stream_ptr()), 'lcgp_nll_grad_host')
        bad = torch.nonzero(self.h_info[:ql])
        if bad.numel():
            k = int(bad[0])
            raise RuntimeError(f'lcgp_b200: Cholesky failed for latent {k % q} of batched emulator {self.active[k // q]} at '
                               f'pivot {int(self.h_info[k])} (NaN/inf inputs)')
        return self.h_out[:E * self.out_len].view(E, self.out_len)

    def __del__(self):
        try:
            for hdl in getattr(self, '_plans', {}).values():
                self.lib.lcgp_plan_destroy(hdl)
        except Exception:
            pass


class _LockStep:
    """Host side of a lock-step batch: the unconstrained variables of E emulators as stacked leaves, the soft-clip
    chain rule and the diag_error_structure expansion done once per step for all of them (torch autograd, same
    element-wise operations as LCGP._get_param_graph, so the values are those of the one-emulator path)."""

    def __init__(self, models: Sequence[LCGP], engine: BatchedEngine):
        self.models, self.eng = list(models), engine
        m0 = models[0]
        self.tf = (m0.lLmb.transform, m0.lLmb0.transform, m0.lnugGPs.transform)
        reps = torch.as_tensor(m0.diag_error_structure, dtype=torch.long)
        self.err_index = torch.repeat_interleave(torch.arange(len(m0.diag_error_structure)), reps)
        self.shapes = [tuple(v.shape) for v in m0.trainable_variables]          # lLmb, lLmb0, lnugGPs, lsigma2s
        self.sizes = [int(np.prod(s)) for s in self.shapes]

    def flat0(self):
        return [m._flat_get() for m in self.models]

    def evaluate(self, xs):
        """xs: E flat unconstrained vectors -> (f (E,), g (E x nvar)) as numpy."""
        E = len(xs)
        U = torch.as_tensor(np.stack(xs), dtype=DT)
        leaves, o = [], 0
        for shp, sz in zip(self.shapes, self.sizes):
            leaves.append(U[:, o:o + sz].reshape((E,) + shp).clone().requires_grad_(True))
            o += sz
        uL, u0, un, us = leaves
        lLmb, lLmb0, lnug = self.tf[0].forward(uL), self.tf[1].forward(u0), self.tf[2].forward(un)
        lsig_p = us[:, self.err_index]
        out = self.eng.evaluate(lLmb.detach(), lLmb0.detach(), lnug.detach(), lsig_p.detach(), True)
        q, d, p = self.eng.q, self.eng.d, self.eng.p
        o1 = 1 + p
        f = out[:, 0].clone().numpy()
        g_sig = out[:, 1:o1]
        g_L = out[:, o1:o1 + q * d].reshape(E, q, d)
        g_0 = out[:, o1 + q * d:o1 + q * d + q]
        g_n = out[:, o1 + q * d + q:o1 + q * d + 2 * q]
        torch.autograd.backward([lLmb, lLmb0, lsig_p, lnug], [g_L.clone(), g_0.clone(), g_sig.clone(), g_n.clone()])
        g = torch.cat([v.grad.reshape(E, -1) for v in leaves], dim=1).numpy()
        return f, g


def _fit_lockstep(models, fit_options, device, compact=True):
    """Fits the models (identical shapes) in lock-step; returns per-model dicts (nfev, nit, success, loss)."""
    opts = {k: v for k, v in (fit_options or {}).items() if k in ('maxcor', 'ftol', 'gtol', 'maxfun', 'maxiter', 'maxls')}
    active = list(range(len(models)))
    stats = [None] * len(models)
    eng = BatchedEngine(models, device)
    ls = _LockStep(models, eng)
    machines = [LbfgsbMachine(x0, **opts) for x0 in ls.flat0()]
    n_steps = 0
    trace = [] if os.environ.get('LCGP_BATCH_TRACE') else None      # (active, engine size, host ms, device+chain-rule ms)
    while True:
        t_a = time.perf_counter()
        for i in list(active):
            if not machines[i].advance():
                active.remove(i)
                mm = machines[i]
                models[i]._flat_set(mm.x)
                stats[i] = dict(nfev=mm.nfev, nit=mm.nit, success=mm.success, message=mm.message, loss=float(mm.f))
        if not active:
            break
        if eng.active != active and (compact or len(active) <= len(eng.active) // 2):
            eng.set_active(active)                        # converged emulators leave the batch
        xs = [machines[i].x for i in eng.active]
        t_b = time.perf_counter()
        f, g = ls.evaluate(xs)
        n_steps += 1
        if trace is not None:
            trace.append((len(active), len(eng.active), 1e3 * (t_b - t_a), 1e3 * (time.perf_counter() - t_b)))
        for slot, i in enumerate(eng.active):
            if not machines[i].done:
                machines[i].supply(f[slot], g[slot])
    for m, s in zip(models, stats):
        m.n_evals = s['nfev']
        m._invalidate_aux()
    if trace:
        tr = np.array(trace)
        print(f'[lcgp_b200.batched] {n_steps} lock-step evaluations; host (L-BFGS-B machines) {tr[:, 2].sum():.0f} ms, '
              f'evaluation (device + chain rule) {tr[:, 3].sum():.0f} ms', flush=True)
        for lo, hi in ((1, 1), (2, 4), (5, 8), (9, 16), (17, 32), (33, 64), (65, 10 ** 9)):
            sel = tr[(tr[:, 1] >= lo) & (tr[:, 1] <= hi)]
            if len(sel):
                print(f'    engine size {lo}-{hi}: {len(sel)} steps, mean active {sel[:, 0].mean():.1f}, host {sel[:, 2].mean():.2f} ms, '
                      f'evaluation {sel[:, 3].mean():.2f} ms per step', flush=True)
    return stats, n_steps


def _same_shape(models):
    m0 = models[0]
    key = lambda m: (int(m.n), int(m.d), int(m.p), int(m.q), m.submethod, tuple(m.diag_error_structure),
                     bool(m.rep_standardize_ybar))
    return all(key(m) == key(m0) for m in models)


def fit_emulators(datasets: Sequence[tuple], model_kwargs: dict | Sequence[dict], optimizer: str = 'L-BFGS-B',
                  fit_options: dict | None = None, threads_per_gpu: int = 4, init_hooks=None, device=None,
                  return_models: bool = False, engine: str = 'auto'):
    """Fit len(datasets) independent LCGP emulators.

    datasets      sequence of (x, y) pairs (x: N x d, y: p x N); pass the same pair several times together
                  with `init_hooks` (e.g. perturbed_restart(seed)) for multi-start fitting
    model_kwargs  one dict for all emulators or one per emulator (q, submethod, ...)
    engine        'lockstep' (one batched device call per step for all of a rank's emulators; needs identical
                  shapes and optimizer='L-BFGS-B'), 'threads' (one host thread + stream + CUDA graph per emulator),
                  'auto' = lockstep whenever it applies
    returns       list (global emulator order, identical on every rank) of dicts with the fitted
                  constrained parameters, final loss, number of evaluations and wall seconds;
                  with return_models=True also the locally fitted LCGP objects (index -> model).
    """
    world, rank = 1, 0
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        world, rank = torch.distributed.get_world_size(), torch.distributed.get_rank()
    n_em = len(datasets)
    mine = list(range(rank, n_em, world))
    kw_of = (lambda i: model_kwargs) if isinstance(model_kwargs, dict) else (lambda i: model_kwargs[i])
    results, models, errors = {}, {}, []

    def result_of(i, m, t0, loss=None):
        lLmb, lLmb0, lsig_p, lnug = m.get_param()
        loss = float(m.loss().detach()) if loss is None else loss
        return dict(index=i, rank=rank, loss=loss, n_evals=m.n_evals, wall_s=time.perf_counter() - t0,
                    lLmb=lLmb.numpy(), lLmb0=lLmb0.numpy(), lsigma2s=m.lsigma2s.numpy(), lnugGPs=lnug.numpy())

    prev_threads = torch.get_num_threads()
    lock_ok = False
    if engine in ('auto', 'lockstep') and optimizer == 'L-BFGS-B' and mine:
        # ---- lock-step batch: construct (host preprocessing, one intra-op thread each: small SVDs), then fit together
        t0 = time.perf_counter()
        torch.set_num_threads(1)
        try:
            built = [None] * len(mine)

            def build(slot):
                i = mine[slot]
                x, y = datasets[i]
                m = LCGP(y=y, x=x, shard=False, device=device, stream_groups=1, **kw_of(i))
                if init_hooks is not None and init_hooks[i] is not None:
                    init_hooks[i](m)
                built[slot] = m
            nt = max(1, min(threads_per_gpu, len(mine)))
            ths = [threading.Thread(target=lambda s0=s0: [build(s) for s in range(s0, len(mine), nt)]) for s0 in range(nt)]
            for t in ths:
                t.start()
            for t in ths:
                t.join()
        finally:
            torch.set_num_threads(prev_threads)
        if all(b is not None for b in built) and _same_shape(built):
            lock_ok = True
            stats, _ = _fit_lockstep(built, fit_options, device)
            for i, m, s in zip(mine, built, stats):
                results[i] = result_of(i, m, t0, s['loss'])      # objective at the final iterate, from the batched call
                results[i].update(nit=s['nit'], converged=s['success'])
                if return_models:
                    models[i] = m
        elif engine == 'lockstep':
            raise ValueError('engine="lockstep" needs emulators of identical (n, d, p, q, submethod, error structure)')
    if not lock_ok:
        lock = threading.Lock()
        todo = list(mine)

        def worker():
            stream = torch.cuda.Stream(device=device)
            while True:
                with lock:
                    if not todo:
                        return
                    i = todo.pop(0)
                try:
                    x, y = datasets[i]
                    with torch.cuda.stream(stream):
                        t0 = time.perf_counter()
                        m = LCGP(y=y, x=x, shard=False, device=device, stream_groups=1, **kw_of(i))   # whole emulator on this thread's stream
                        if init_hooks is not None and init_hooks[i] is not None:
                            init_hooks[i](m)
                        m.fit(optimizer=optimizer, **(fit_options or {}))
                        res = result_of(i, m, t0)
                    with lock:
                        results[i] = res
                        if return_models:
                            models[i] = m
                except Exception as ex:   # surface worker failures in the caller
                    with lock:
                        errors.append((i, ex))

        # Host-side preprocessing of a small emulator (sort, SVD of a p x n matrix) is slower with many
        # intra-op threads than with one, and several emulators are preprocessed concurrently anyway.
        torch.set_num_threads(1)
        try:
            threads = [threading.Thread(target=worker) for _ in range(max(1, min(threads_per_gpu, len(mine))))]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
        finally:
            torch.set_num_threads(prev_threads)
        if errors:
            raise RuntimeError(f'fit_emulators: emulator {errors[0][0]} failed: {errors[0][1]!r}') from errors[0][1]

    if world > 1:
        gathered = [None] * world
        torch.distributed.all_gather_object(gathered, results)
        merged = {}
        for part in gathered:
            merged.update(part)
        results = merged
    ordered = [results[i] for i in range(n_em)]
    return (ordered, models) if return_models else ordered
