"""ctypes binding of liblcgp_b200.so (C-ABI declared in include/lcgp_b200.h).

There is no CPU fallback: importing this module never fails (so host-side logic can be tested on a
CPU-only box), but `lib()` raises if the shared library has not been built, and every compute entry
point needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('LCGP_B200_LIB') or os.path.join(_HERE, '_lib', 'liblcgp_b200.so')   # (override: A/B builds)
NB = 128
MAX_D = 64
N_STAGE_EVENTS = 7

_dp = C.c_void_p  # device / host pointers travel as integers


class Problem(C.Structure):
    """struct lcgp_problem."""
    _fields_ = [('n', C.c_int32), ('d', C.c_int32), ('p', C.c_int32), ('q_loc', C.c_int32),
                ('include_host_terms', C.c_int32), ('n_emu', C.c_int32),
                ('scale', C.c_double), ('sum_log_r', C.c_double),
                ('X', _dp), ('sr', _dp), ('YR', _dp), ('w', _dp), ('t', _dp), ('phi', _dp), ('D', _dp),
                ('emu_consts', _dp)]


_SIGS = {
    'lcgp_version': (C.c_char_p, []),
    'lcgp_launch_count': (C.c_ulonglong, []),
    'lcgp_out_len': (C.c_size_t, [C.c_int32] * 3),
    'lcgp_workspace_bytes': (C.c_size_t, [C.c_int32] * 4),
    'lcgp_workspace_bytes_batched': (C.c_size_t, [C.c_int32] * 5),
    'lcgp_predict_scratch_bytes': (C.c_size_t, [C.c_int32] * 3),
    'lcgp_nll_grad': (C.c_int, [C.POINTER(Problem), _dp, _dp, _dp, _dp, _dp, C.c_size_t, _dp, _dp, C.c_int32,
                                C.POINTER(C.c_void_p), _dp]),
    'lcgp_nll_grad_host': (C.c_int, [C.POINTER(Problem), _dp, _dp, _dp, _dp, _dp, C.c_size_t, _dp, _dp, C.c_int32,
                                     C.POINTER(C.c_void_p), _dp]),
    'lcgp_plan_create': (C.c_int, [C.POINTER(Problem), _dp, C.c_size_t, _dp, _dp, _dp, C.c_int32, C.POINTER(C.c_void_p)]),
    'lcgp_plan_run': (C.c_int, [C.c_void_p, _dp]),
    'lcgp_plan_is_graph': (C.c_int, [C.c_void_p]),
    'lcgp_plan_destroy': (None, [C.c_void_p]),
    'lcgp_predict': (C.c_int, [C.POINTER(Problem), _dp, _dp, _dp, _dp, C.c_size_t, _dp, C.c_int32, C.c_int32,
                               _dp, C.c_size_t, _dp, _dp, _dp]),
    'lcgp_predict_outputs': (C.c_int, [_dp, _dp, _dp, _dp, _dp, _dp, C.c_int32, C.c_int32, C.c_int32, _dp, _dp, _dp, _dp]),
    'lcgp_predict_fullcov': (C.c_int, [_dp, _dp, _dp, _dp, C.c_int32, C.c_int32, C.c_int32, _dp, _dp]),
    'lcgp_prep_segment_mean': (C.c_int, [_dp, _dp, _dp, C.c_int32, C.c_int32, C.c_int32, _dp, _dp]),
    'lcgp_prep_row_select': (C.c_int, [_dp, _dp, C.c_int32, C.c_int32, C.c_int32, _dp, _dp]),
    'lcgp_prep_standardize': (C.c_int, [_dp, _dp, _dp, _dp, C.c_int32, C.c_int32, _dp, _dp, _dp, _dp]),
    'lcgp_grad_phi': (C.c_int, [C.POINTER(Problem), _dp, _dp, C.c_size_t, _dp, _dp]),
    'lcgp_pack_sharded': (C.c_int, [_dp, _dp, _dp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _dp, _dp]),
    'lcgp_get_aux': (C.c_int, [C.POINTER(Problem), _dp, C.c_size_t, _dp, _dp, _dp]),
    'lcgp_get_Ainv': (C.c_int, [C.POINTER(Problem), _dp, C.c_size_t, C.c_int32, _dp, _dp]),
    'lcgp_kernel_matrix': (C.c_int, [_dp, C.c_int32, _dp, C.c_int32, C.c_int32, _dp, _dp, _dp, C.c_int32, _dp, _dp]),
    'lcgp_build_A': (C.c_int, [_dp, _dp, C.c_int32, C.c_int32, _dp, _dp, _dp, _dp, C.c_int32, _dp, C.c_int32, _dp]),
    'lcgp_potrf_scratch_bytes': (C.c_size_t, [C.c_int32, C.c_int32]),
    'lcgp_potrf_batched': (C.c_int, [_dp, C.c_int32, C.c_int32, _dp, _dp, _dp, _dp, _dp, C.c_size_t, _dp]),
    'lcgp_potrf_trtri_batched': (C.c_int, [_dp, C.c_int32, C.c_int32, _dp, _dp, _dp, _dp, _dp, C.c_size_t, _dp]),
    'lcgp_trtri_scratch_bytes': (C.c_size_t, [C.c_int32, C.c_int32]),
    'lcgp_trtri_batched': (C.c_int, [_dp, C.c_int32, C.c_int32, _dp, _dp, _dp, C.c_size_t, _dp]),
}
EXPORTS = tuple(_SIGS)

_lib = None


def lib():
    """The loaded shared library; raises RuntimeError if it was never built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'lcgp_b200: CUDA library not built ({LIB_PATH} missing). Run '
                f'`python -c "import __graft_entry__ as g; g.build()"` or `make -C lcgp_b200/csrc`. '
                f'There is no CPU fallback.')
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class LCGPError(RuntimeError):
    pass


def check(rc: int, what: str):
    if rc == 0:
        return
    if rc >= 1000:
        raise LCGPError(f'{what}: CUDA error {rc - 1000}')
    names = {-1: 'invalid argument', -2: 'unsupported dimension (d > 64 or bad padding)', -3: 'workspace too small'}
    raise LCGPError(f'{what}: {names.get(rc, rc)}')


def available() -> bool:
    """True when the shared library has been built (says nothing about a CUDA device)."""
    return os.path.exists(LIB_PATH)


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError('lcgp_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.')
    lib()


def ptr(t):
    """Device/host pointer of a contiguous fp64/int32 torch tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_contiguous()
    return t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


def padded(n: int) -> int:
    return (n + NB - 1) // NB * NB
