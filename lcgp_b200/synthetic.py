"""Synthetic data for the BASELINE.json configs.

The replicated 1-d generators reproduce the random streams of the reference's
illustration (illustration-examples/lcgp-rep-3d-illustration.py:13-104, also cell 8 of
lcgp-rep-1d-illustration.ipynb) so that the notebook's stored outputs can serve as golden
vectors; the large configs are new.
"""
from __future__ import annotations

import numpy as np


def f_true(x):
    """Three noise-free output curves (lcgp-rep-3d-illustration.py:13-18)."""
    x = np.asarray(x, dtype=np.float64)
    return np.vstack([0.8 + 0.3 * np.sin(2 * np.pi * x) + 0.2 * x,
                      0.3 + 0.5 * np.cos(2 * np.pi * x),
                      -0.4 - (x - 0.5) ** 2 + 0.2 * np.sin(4 * np.pi * x)])


def _emit(x_unique, reps, rng, noise_std):
    xs, ys = [], []
    for xi, r in zip(x_unique, reps):
        base = f_true([xi])[:, 0]
        for _ in range(int(r)):
            # one normal draw per output, in output order (same stream as the reference)
            eps = np.array([rng.normal(0, s) for s in noise_std], dtype=np.float64)
            xs.append([xi])
            ys.append(base + eps)
    xtest = np.linspace(0.0, 1.0, 400, dtype=np.float64)[:, None]
    return (np.array(xs, dtype=np.float64), np.array(ys, dtype=np.float64).T,
            xtest, f_true(xtest[:, 0]))


def rep1d_uniform(n_unique=16, rep_choices=(1, 2, 3, 4, 5), noise_std=(0.05, 0.08, 0.10), seed=2025):
    """Case 1 (lcgp-rep-3d-illustration.py:20-44, 106-113): all counts drawn first."""
    rng = np.random.default_rng(seed)
    x_unique = np.linspace(0.0, 1.0, n_unique, dtype=np.float64)
    reps = rng.choice(rep_choices, size=n_unique, replace=True)
    return _emit(x_unique, reps, rng, noise_std)


def rep1d_skewed(n_unique=40, heavy_region=(0.20, 0.45), light_rep_choices=(1, 2),
                 heavy_rep_choices=(8, 12, 16, 20), noise_std=(0.05, 0.08, 0.10), seed=123):
    """Case 2 (lcgp-rep-3d-illustration.py:46-71, 116-125): count drawn per location, interleaved
    with that location's noise draws.  seed=123 gives the notebook's N=194, n=40."""
    rng = np.random.default_rng(seed)
    x_unique = np.linspace(0.0, 1.0, n_unique, dtype=np.float64)
    xs, ys = [], []
    for xi in x_unique:
        heavy = heavy_region[0] <= xi <= heavy_region[1]
        r = int(rng.choice(heavy_rep_choices) if heavy else rng.choice(light_rep_choices))
        base = f_true([xi])[:, 0]
        for _ in range(r):
            eps = np.array([rng.normal(0, s) for s in noise_std], dtype=np.float64)
            xs.append([xi])
            ys.append(base + eps)
    xtest = np.linspace(0.0, 1.0, 400, dtype=np.float64)[:, None]
    return (np.array(xs, dtype=np.float64), np.array(ys, dtype=np.float64).T,
            xtest, f_true(xtest[:, 0]))


def rep1d_hotspots(n_unique=50, hotspots=((0.15, 10, 15), (0.50, 18, 25), (0.80, 12, 20)),
                   base_rep_choices=(1,), noise_std=(0.05, 0.08, 0.10), seed=7):
    """Case 3 (lcgp-rep-3d-illustration.py:73-104, 128-136)."""
    rng = np.random.default_rng(seed)
    x_unique = np.linspace(0.0, 1.0, n_unique, dtype=np.float64)
    hot = {int(np.argmin(np.abs(x_unique - x0))): (lo, hi) for (x0, lo, hi) in hotspots}
    xs, ys = [], []
    for i, xi in enumerate(x_unique):
        if i in hot:
            r = int(rng.integers(hot[i][0], hot[i][1] + 1))
        else:
            r = int(rng.choice(base_rep_choices))
        base = f_true([xi])[:, 0]
        for _ in range(r):
            eps = np.array([rng.normal(0, s) for s in noise_std], dtype=np.float64)
            xs.append([xi])
            ys.append(base + eps)
    xtest = np.linspace(0.0, 1.0, 400, dtype=np.float64)[:, None]
    return (np.array(xs, dtype=np.float64), np.array(ys, dtype=np.float64).T,
            xtest, f_true(xtest[:, 0]))


def rep3d(seed=3, n_unique=200, d=3, p=3, rep_choices=(1, 2, 3, 4), noise=0.1):
    """True 3-d-input replicated case for config 2: y = sin(xW)^T + noise."""
    rng = np.random.default_rng(seed)
    xu = rng.uniform(0, 1, (n_unique, d))
    r = rng.choice(rep_choices, size=n_unique)
    x = np.repeat(xu, r, axis=0)
    W = rng.standard_normal((d, p))
    y = np.sin(x @ W).T + noise * rng.standard_normal((p, x.shape[0]))
    x0 = rng.uniform(0, 1, (64, d))
    return x, y, x0


def latent_mixture(n, d, p, q_true, seed, rep_choices=None, noise=0.05, n0=256):
    """Configs 3-5: X ~ U(0,1)^{n x d}; latent truth g_k(x) = sum_j sin(2 pi f_kj x_j
    + phase_kj), f ~ U(0.5, 2); y = W g + noise * eps with W ~ N(0,1)^{p x q_true}.
    rep_choices=None gives unreplicated data (submethod='full'); otherwise each unique row is
    repeated r_i in rep_choices times (submethod='rep')."""
    rng = np.random.default_rng(seed)
    xu = rng.uniform(0, 1, (n, d))
    f = rng.uniform(0.5, 2.0, (q_true, d))
    ph = rng.uniform(0, 2 * np.pi, (q_true, d))
    W = rng.standard_normal((p, q_true))

    def truth(xx):
        g = np.sin(2 * np.pi * xx[:, None, :] * f[None] + ph[None]).sum(axis=2)   # (m, q_true)
        return W @ g.T                                                            # (p, m)

    if rep_choices is None:
        x = xu
    else:
        r = rng.choice(rep_choices, size=n)
        x = np.repeat(xu, r, axis=0)
    y = truth(x) + noise * rng.standard_normal((p, x.shape[0]))
    x0 = rng.uniform(0, 1, (n0, d))
    return x, y, x0, truth(x0)


CONFIGS = {
    # name: kwargs for latent_mixture + model kwargs (BASELINE.json configs 3, 4, 5)
    'cfg3_full': dict(data=dict(n=2000, d=8, p=500, q_true=10, seed=2000, rep_choices=None),
                      model=dict(q=10, submethod='full')),
    'cfg3_rep': dict(data=dict(n=2000, d=8, p=500, q_true=10, seed=2000, rep_choices=(1, 2, 3, 4)),
                     model=dict(q=10, submethod='rep')),
    'cfg4_rep': dict(data=dict(n=8000, d=10, p=2000, q_true=32, seed=8000, rep_choices=(1, 2, 3)),
                     model=dict(q=32, submethod='rep')),
    # per-GPU share of config 4 at N = 8 (4 of the 32 latents), for single-GPU tuning of the sharded case
    'cfg4_shard8': dict(data=dict(n=8000, d=10, p=2000, q_true=32, seed=8000, rep_choices=(1, 2, 3)),
                        model=dict(q=4, submethod='rep')),
    'cfg5_one': dict(data=dict(n=1024, d=6, p=64, q_true=8, seed=1024, rep_choices=None),
                     model=dict(q=8, submethod='full')),
}


def make_config(name, **override):
    cfg = CONFIGS[name]
    kw = dict(cfg['data'])
    kw.update(override)
    x, y, x0, y0 = latent_mixture(**kw)
    return x, y, x0, y0, dict(cfg['model'])
