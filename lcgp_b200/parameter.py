"""Constrained parameters of the LCGP model (host side).

The reference keeps four `gpflow.Parameter`s, three of them behind a TFP `SoftClip` bijector
(lcgp.py:181-211); the optimizer works on the unconstrained values.  This module provides the
same surface on torch: a two-sided soft clip and a small `Parameter` holder with `.numpy()`,
`.assign()` and an `.unconstrained` leaf tensor.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

DT = torch.float64


def _inv_softplus(a):
    return a + torch.log(-torch.expm1(-a))


class SoftClip:
    """y = high - softplus(high - low - softplus(u - low)) * (high - low) / softplus(high - low)
    (TFP bijectors.SoftClip, hinge_softness 1).  Maps R onto (low, high)."""

    def __init__(self, low: float, high: float):
        self.low = float(low)
        self.high = float(high)
        w = torch.tensor(self.high - self.low, dtype=DT)
        self._w = w
        self._c = w / F.softplus(w)

    def forward(self, u):
        return self.high - F.softplus(self._w - F.softplus(u - self.low)) * self._c

    def inverse(self, y):
        y = torch.as_tensor(y, dtype=DT)
        return self.low + _inv_softplus(self._w - _inv_softplus((self.high - y) / self._c))


class Identity:
    low, high = -math.inf, math.inf

    def forward(self, u):
        return u

    def inverse(self, y):
        return torch.as_tensor(y, dtype=DT)


class Parameter:
    """Holder of one trainable array: `.unconstrained` is the leaf the optimizer updates,
    `.value()` the constrained tensor (differentiable), `.numpy()` / `.assign()` act on the
    constrained value like gpflow.Parameter."""

    def __init__(self, value, transform=None, name=''):
        self.transform = transform or Identity()
        self.name = name
        self.unconstrained = self.transform.inverse(torch.as_tensor(np.asarray(value), dtype=DT)) \
            .clone().detach().requires_grad_(True)

    def value(self):
        return self.transform.forward(self.unconstrained)

    def assign(self, value):
        v = torch.as_tensor(np.asarray(value, dtype=np.float64), dtype=DT).reshape(self.unconstrained.shape)
        with torch.no_grad():
            self.unconstrained.copy_(self.transform.inverse(v))
        return self

    def numpy(self):
        return self.value().detach().numpy()

    def detach(self):
        return self.value().detach()

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a.astype(dtype) if dtype is not None else a

    def __getitem__(self, idx):
        return self.value().detach()[idx]

    @property
    def shape(self):
        return self.unconstrained.shape

    def __repr__(self):
        return f'Parameter({self.name!r}, shape={tuple(self.shape)}, bounds=({self.transform.low}, {self.transform.high}))'
