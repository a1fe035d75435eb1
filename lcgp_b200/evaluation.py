"""Prediction diagnostics with the reference's definitions (src/lcgp/evaluation.py:5-63), numpy only."""
import numpy as np
from scipy.stats import norm


def rmse(y, ypredmean):
    return np.sqrt(np.mean((y - ypredmean) ** 2))


def normalized_rmse(y, ypredmean):
    """Errors are divided by each output's range before averaging."""
    span = np.ptp(y, axis=1).reshape(-1, 1)
    return np.sqrt(np.mean(((y - ypredmean) / span) ** 2))


def dss(y, ypredmean, ypredcov, use_diag):
    """Dawid-Sebastiani score averaged over test points; ypredcov is (p, n) variances when
    use_diag, else (p, p, n) covariances."""
    resid = y - ypredmean
    n = y.shape[1]
    if use_diag:
        return (np.log(ypredcov).sum() + (resid ** 2 / ypredcov).sum()) / n
    total = 0.0
    for i in range(n):
        S = ypredcov[:, :, i]
        w, U = np.linalg.eigh(S)
        z = (resid[:, i] @ U) / np.sqrt(w)
        total += np.linalg.slogdet(S)[1] + (z ** 2).sum()
    return total / n


def intervalstats(y, ypredmean, ypredvar):
    """Empirical coverage and mean width of the central 95 % predictive interval."""
    half = np.sqrt(ypredvar) * norm.ppf(0.975)
    lo, hi = ypredmean - half, ypredmean + half
    return np.mean((y >= lo) & (y <= hi)), np.mean(hi - lo)
