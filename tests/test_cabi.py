"""The C-ABI shared library: builds for sm_100a, loads without a GPU, and exports exactly the entry
points include/lcgp_b200.h declares.  No compute call is made here."""
import ctypes
import os
import re

from lcgp_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, 'include', 'lcgp_b200.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(lcgp_[a-zA-Z0-9_]+)\s*\(', txt)))


def test_build_and_exports():
    import __graft_entry__ as g
    g.build()
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    decl = _declared()
    assert len(decl) >= 14
    for name in decl:
        assert hasattr(lib, name), f'{name} declared in include/lcgp_b200.h but not exported'
    assert sorted(_cabi.EXPORTS) == decl, 'ctypes prototypes and header are out of sync'


def test_size_queries_without_gpu():
    L = _cabi.lib()
    assert b'sm_100a' in L.lcgp_version()
    assert L.lcgp_out_len(2000, 10, 32) == 1 + 2000 + 320 + 64 + 64
    b4 = L.lcgp_workspace_bytes(8000, 10, 2000, 4)
    b32 = L.lcgp_workspace_bytes(8000, 10, 2000, 32)
    assert 4 * 8064 * 8064 * 8 < b4 < 1.5 * 4 * 8064 * 8064 * 8          # factor buffers dominate
    assert 7.5 < b32 / b4 < 8.5
    assert L.lcgp_workspace_bytes(0, 10, 2000, 4) == 0
    assert L.lcgp_predict_scratch_bytes(8000, 4, 256) >= 4 * 256 * 8064 * 8
    assert L.lcgp_trtri_scratch_bytes(128, 3) == 0 and L.lcgp_trtri_scratch_bytes(129, 3) == 0


def test_problem_struct_layout():
    # the ctypes mirror must match the C struct: 6 int32, 2 doubles, 7 pointers, no padding surprises
    assert ctypes.sizeof(_cabi.Problem) == 6 * 4 + 2 * 8 + 8 * 8
    assert _cabi.Problem.scale.offset == 24 and _cabi.Problem.X.offset == 40


def test_sass_has_fp64_tensor_and_async_copy():
    """Evidence that the hot kernels are the sm_100a FP64 tensor path (DMMA) fed by cp.async."""
    import shutil
    import subprocess
    if shutil.which('cuobjdump') is None:
        import pytest
        pytest.skip('cuobjdump not available')
    sass = subprocess.run(['cuobjdump', '-sass', _cabi.LIB_PATH], capture_output=True, text=True).stdout
    assert 'sm_100a' in sass
    assert sass.count('DMMA.8x8x4') >= 128                # FP64 tensor pipe
    assert sass.count('UTMALDG') >= 100                    # TMA tensor-tile loads: the DEFAULT staging engine
    assert 'SYNCS' in sass                                 # mbarriers (TMA completion / full-empty ring)
    assert 'LDGSTS' in sass                                # cp.async: the fallback engine
    # the persistent Cholesky kernel itself: DMMA + TMA + global-memory dependency flags in one kernel
    pll = sass[sass.index('potrf_pll_kernel'):]
    pll = pll[:pll.index('Function :', 20) if 'Function :' in pll[20:] else len(pll)]
    assert 'DMMA' in pll and 'UTMALDG' in pll and 'NANOSLEEP' in pll
