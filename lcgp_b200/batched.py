"""Batched fitting of independent emulators / L-BFGS restarts (BASELINE.json config 5).

The emulators are independent, so there is no data-path collective ("replicas only"):
emulator i is fitted by rank `i mod world` on that rank's GPU, and only the fitted parameters are
gathered at the end.  Inside a rank several host threads drive one emulator each on its own CUDA
stream: an n = 1024 evaluation is a latency-bound chain of small kernels that cannot fill 148 SMs,
and the C-ABI call releases the GIL, so concurrent emulators fill the GPU instead.
"""
from __future__ import annotations

import threading
import time
from typing import Callable, Sequence

import numpy as np
import torch

from .model import LCGP


def perturbed_restart(seed: int, lo: float = 0.5, hi: float = 2.0) -> Callable[[LCGP], None]:
    """Start-point perturbation for restarts: length-scales multiplied by U(lo, hi) factors."""
    def apply(model: LCGP):
        rng = np.random.default_rng(seed)
        model.lLmb.assign(model.lLmb.numpy() * rng.uniform(lo, hi, model.lLmb.numpy().shape))
    return apply


def fit_emulators(datasets: Sequence[tuple], model_kwargs: dict | Sequence[dict], optimizer: str = 'L-BFGS-B',
                  fit_options: dict | None = None, threads_per_gpu: int = 4, init_hooks=None, device=None,
                  return_models: bool = False):
    """Fit len(datasets) independent LCGP emulators.

    datasets      sequence of (x, y) pairs (x: N x d, y: p x N); pass the same pair several times together
                  with `init_hooks` (e.g. perturbed_restart(seed)) for multi-start fitting
    model_kwargs  one dict for all emulators or one per emulator (q, submethod, ...)
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
                    loss = float(m.loss())
                    lLmb, lLmb0, lsig_p, lnug = m.get_param()
                    res = dict(index=i, rank=rank, loss=loss, n_evals=m.n_evals, wall_s=time.perf_counter() - t0,
                               lLmb=lLmb.numpy(), lLmb0=lLmb0.numpy(), lsigma2s=m.lsigma2s.numpy(), lnugGPs=lnug.numpy())
                with lock:
                    results[i] = res
                    if return_models:
                        models[i] = m
            except Exception as ex:   # surface worker failures in the caller
                with lock:
                    errors.append((i, ex))

    # Host-side preprocessing of a small emulator (sort, SVD of a p x n matrix) is slower with many
    # intra-op threads than with one, and several emulators are preprocessed concurrently anyway.
    prev_threads = torch.get_num_threads()
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
