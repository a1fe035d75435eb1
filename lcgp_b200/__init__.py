"""lcgp_b200 -- B200-native (sm_100a) implementation of LCGP's emulator-fitting hot path.

Public surface mirrors the reference package (`from lcgp import LCGP, Matern32`):
    LCGP      model object (constructor, fit, predict, loss / neglpost / neglpost_rep, get_param)
    Matern32  covariance operator
    evaluation  rmse / normalized_rmse / dss / intervalstats
"""
from .model import LCGP  # noqa: F401
from .kernels import Matern32  # noqa: F401
from . import evaluation  # noqa: F401
from .batched import fit_emulators, perturbed_restart  # noqa: F401
from .harness import LCGPRun, SuperRun  # noqa: F401

__version__ = '0.1.0'
