"""`Matern32` operator -- the reference's covariance entry point (src/lcgp/covmat.py:5-55) on sm_100a.

Same signature, same argument meaning (`llmb` is the per-dimension length-scale itself, `llmb0`
the variance multiplier, `lnug` the raw nugget odds, nug = lnug / (1 + lnug)), same assertion
behaviour.  The matrix is produced by `lcgp_kernel_matrix` (csrc/matern.cu) on the current CUDA
device and returned on the device of `x1`.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _cabi

DT = torch.float64


def _as2d_check(x, name):
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(np.asarray(x), dtype=DT)
    assert x.ndim == 2, f'input {name} should be 2-dimensional, (n_param, dim_param)'
    return x.to(DT)


def Matern32(x1, x2, llmb, llmb0, lnug, diag_only: bool = False):
    x1 = _as2d_check(x1, 'x1')
    x2 = _as2d_check(x2, 'x2')
    assert x1.shape[1] == x2.shape[1], 'the dim_param of input x1 and x2 should be the same.'
    ret_dev = x1.device
    llmb0_t = torch.as_tensor(llmb0, dtype=DT).reshape(-1)[:1]

    if diag_only:   # covmat.py:23-29
        assert bool(torch.all(torch.abs(x1 - x2.to(x1.device)) <= (1e-6 + 1e-6 * torch.abs(x2.to(x1.device))))), \
            'diag_only should only be called when x1 and x2 are identical.'
        return llmb0_t.to(ret_dev) * torch.ones(x1.shape[0], dtype=DT, device=ret_dev)

    _cabi.require_cuda()
    dev = torch.device('cuda', torch.cuda.current_device())
    a = x1.detach().to(dev).contiguous()
    b = x2.detach().to(dev).contiguous()
    ell = torch.as_tensor(llmb, dtype=DT).detach().reshape(-1).to(dev).contiguous()
    assert ell.numel() == a.shape[1], 'llmb needs one length-scale per input dimension'
    s0 = llmb0_t.detach().to(dev).contiguous()
    nug = torch.as_tensor(lnug, dtype=DT).detach().reshape(-1)[:1].to(dev).contiguous()
    same = int(a.shape == b.shape and bool(torch.equal(a, b)))   # covmat.py:46-49
    out = torch.empty((a.shape[0], b.shape[0]), dtype=DT, device=dev)
    rc = _cabi.lib().lcgp_kernel_matrix(a.data_ptr(), a.shape[0], b.data_ptr(), b.shape[0], a.shape[1],
                                        ell.data_ptr(), s0.data_ptr(), nug.data_ptr(), same, out.data_ptr(),
                                        _cabi.stream_ptr())
    _cabi.check(rc, 'lcgp_kernel_matrix')
    return out.to(ret_dev)
