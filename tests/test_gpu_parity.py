"""Parity of the CUDA path (through the C-ABI, as lcgp_b200.LCGP / Matern32 call it) against the CPU
oracle on identical inputs.  Tolerances are BASELINE.json's: objective 1e-10 relative, gradients 1e-8,
fitted hyper-parameters and predictions 1e-6."""
import json
import os

import numpy as np
import pytest
import torch

from lcgp_b200 import LCGP, Matern32, _cabi, evaluation, synthetic
from oracle import lcgp_oracle as O
from helpers import make_full_data, make_ragged_rep_data, make_rep_data, move_params

pytestmark = pytest.mark.gpu
DT = torch.float64
NLL_TOL, GRAD_TOL, PRED_TOL = 1e-10, 1e-8, 1e-6
HERE = os.path.dirname(__file__)
FIX = json.load(open(os.path.join(HERE, 'golden', 'oracle_fixtures.json')))
GOLD = json.load(open(os.path.join(HERE, 'golden', 'notebook_case2.json')))


def rel(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


# ---------------------------------------------------------------- a1: kernel matrix
@pytest.mark.parametrize('n1,n2,d', [(1, 1, 1), (7, 5, 1), (64, 64, 2), (65, 130, 3), (300, 300, 8), (129, 40, 10)])
def test_matern32_matches_oracle(n1, n2, d):
    rng = np.random.default_rng(n1 * 1000 + n2)
    x1 = torch.as_tensor(rng.uniform(0, 1, (n1, d)))
    x2 = x1.clone() if n1 == n2 else torch.as_tensor(rng.uniform(0, 1, (n2, d)))
    ell = torch.as_tensor(rng.uniform(0.2, 2.0, d)); s0 = torch.tensor(1.7, dtype=DT); nug = torch.tensor(3e-4, dtype=DT)
    C = Matern32(x1, x2, ell, s0, nug)
    Co = O.Matern32(x1, x2, ell, s0, nug)
    assert C.shape == (n1, n2) and C.device.type == 'cpu'
    assert rel(C, Co) < 1e-13                          # nugget only when the point sets are identical
    if n1 == n2 and n1 > 1:
        x2b = x2.clone(); x2b[0, 0] += 1e-3             # same shape, different values: no nugget
        assert rel(Matern32(x1, x2b, ell, s0, nug), O.Matern32(x1, x2b, ell, s0, nug)) < 1e-13


def test_matern32_argument_checks():                    # test_cov.py:18-23, covmat.py:18-29
    x = torch.rand(5, 2, dtype=DT)
    with pytest.raises(AssertionError):
        Matern32(torch.rand(5, dtype=DT), x, torch.ones(2), 1.0, 1e-3)
    with pytest.raises(AssertionError):
        Matern32(x, torch.rand(5, 3, dtype=DT), torch.ones(2), 1.0, 1e-3)
    with pytest.raises(AssertionError):
        Matern32(x, x + 1.0, torch.ones(2), 1.0, 1e-3, diag_only=True)
    assert torch.equal(Matern32(x, x, torch.ones(2), 2.5, 1e-3, diag_only=True), 2.5 * torch.ones(5, dtype=DT))


# ---------------------------------------------------------------- stages through the C-ABI
def _potrf(L, F, npad, q, DL, DU, ldp, info, st, persistent=True):
    """lcgp_potrf_batched with (persistent left-looking kernel) or without (launch chain) the flag scratch;
    persistent='fused': lcgp_potrf_trtri_batched (factor and triangular inverse in one launch)."""
    if persistent == 'fused':
        sb = int(L.lcgp_potrf_scratch_bytes(npad, q))
        scr = torch.full((sb // 4,), 0x7f7f7f7f, dtype=torch.int32, device=F.device)
        rc = L.lcgp_potrf_trtri_batched(F.data_ptr(), npad, q, DL.data_ptr(), DU.data_ptr(), ldp, info.data_ptr(), scr.data_ptr(), sb, st)
        torch.cuda.synchronize()
        return rc
    if not persistent:
        return L.lcgp_potrf_batched(F.data_ptr(), npad, q, DL.data_ptr(), DU.data_ptr(), ldp, info.data_ptr(), None, 0, st)
    sb = int(L.lcgp_potrf_scratch_bytes(npad, q))
    scr = torch.full((sb // 4,), 0x7f7f7f7f, dtype=torch.int32, device=F.device)     # garbage: the call must clear it
    rc = L.lcgp_potrf_batched(F.data_ptr(), npad, q, DL.data_ptr(), DU.data_ptr(), ldp, info.data_ptr(), scr.data_ptr(), sb, st)
    torch.cuda.synchronize()
    return rc


@pytest.mark.parametrize('persistent', [True, False, 'fused'])
@pytest.mark.parametrize('n,d,q', [(40, 1, 2), (128, 2, 1), (129, 3, 2), (700, 5, 3), (1500, 8, 2), (2100, 4, 9)])
def test_build_potrf_trtri_stages(n, d, q, persistent):
    L = _cabi.lib()
    dev = torch.device('cuda')
    rng = np.random.default_rng(n)
    X = torch.as_tensor(rng.uniform(0, 1, (n, d))); r = torch.as_tensor(rng.integers(1, 4, n).astype(float))
    sr = torch.sqrt(r)
    ell = torch.as_tensor(rng.uniform(0.3, 1.5, (q, d))); s0 = torch.as_tensor(rng.uniform(0.5, 3, q))
    nug = torch.as_tensor(np.exp(rng.uniform(-12, -5, q))); D = torch.as_tensor(rng.uniform(0.3, 2, q))
    npad = _cabi.padded(n); nb = npad // 128; st = _cabi.stream_ptr()
    g = lambda t: t.to(dev).contiguous()
    Xd, srd, elld, s0d, nugd, Dd = map(g, (X, sr, ell, s0, nug, D))
    F = torch.full((q, npad, npad), float('nan'), dtype=DT, device=dev)
    assert L.lcgp_build_A(Xd.data_ptr(), srd.data_ptr(), n, d, elld.data_ptr(), s0d.data_ptr(), nugd.data_ptr(),
                          Dd.data_ptr(), q, F.data_ptr(), npad, st) == 0
    A = torch.stack([torch.eye(npad, dtype=DT) for _ in range(q)])
    for k in range(q):
        Ck = O.Matern32(X, X, ell[k], s0[k], nug[k])
        A[k, :n, :n] = torch.eye(n, dtype=DT) + D[k] * ((Ck * sr[None, :]) * sr[:, None])     # lcgp.py:616
    low = torch.tril(torch.ones(npad, npad, dtype=torch.bool))
    assert rel(F.cpu()[:, low], A[:, low]) < 1e-14
    DL = torch.zeros((q, nb, 128, 128), dtype=DT, device=dev); DU = torch.zeros_like(DL)
    ldp = torch.zeros((q, nb), dtype=DT, device=dev); info = torch.ones(q, dtype=torch.int32, device=dev)
    assert _potrf(L, F, npad, q, DL, DU, ldp.data_ptr(), info, st, persistent) == 0
    Lref = torch.linalg.cholesky(A)
    assert info.cpu().tolist() == [0] * q
    assert rel(F.cpu()[:, low], Lref[:, low]) < 1e-12
    assert rel(ldp.sum(1), torch.log(torch.diagonal(Lref, dim1=1, dim2=2)).sum(1)) < 1e-12
    if persistent != 'fused':
        sb = int(L.lcgp_trtri_scratch_bytes(npad, q))
        scr = torch.empty(max(sb // 8, 1), dtype=DT, device=dev)
        assert L.lcgp_trtri_batched(F.data_ptr(), npad, q, DL.data_ptr(), DU.data_ptr(), scr.data_ptr(), sb, st) == 0
    Uref = torch.linalg.solve_triangular(Lref, torch.eye(npad, dtype=DT).expand_as(Lref), upper=False).transpose(1, 2)
    Fc = F.cpu()
    for I in range(nb):
        sl = slice(I * 128, (I + 1) * 128)
        assert rel(DU[:, I], Uref[:, sl, sl]) < 1e-11 and rel(DL[:, I], Uref[:, sl, sl].transpose(1, 2)) < 1e-11
        if I + 1 < nb:
            assert rel(Fc[:, sl, (I + 1) * 128:], Uref[:, sl, (I + 1) * 128:]) < 1e-11


@pytest.mark.parametrize('persistent', [True, False])
def test_potrf_reports_bad_pivot(persistent):
    L = _cabi.lib(); dev = torch.device('cuda')
    F = torch.eye(256, dtype=DT, device=dev).repeat(2, 1, 1).contiguous()
    F[1, 130, 130] = -1.0
    DL = torch.zeros((2, 2, 128, 128), dtype=DT, device=dev); DU = torch.zeros_like(DL)
    info = torch.zeros(2, dtype=torch.int32, device=dev)
    assert _potrf(L, F, 256, 2, DL, DU, None, info, _cabi.stream_ptr(), persistent) == 0
    assert info.cpu().tolist() == [0, 131]
    assert L.lcgp_potrf_batched(F.data_ptr(), 200, 2, DL.data_ptr(), DU.data_ptr(), None, info.data_ptr(), None, 0, _cabi.stream_ptr()) == -2
    assert L.lcgp_potrf_batched(None, 256, 2, DL.data_ptr(), DU.data_ptr(), None, info.data_ptr(), None, 0, _cabi.stream_ptr()) == -1
    assert L.lcgp_potrf_batched(F.data_ptr(), 256, 2, DL.data_ptr(), DU.data_ptr(), None, info.data_ptr(), F.data_ptr(), 8, _cabi.stream_ptr()) == -3


# ---------------------------------------------------------------- a2-a4: objective and gradient
def _pair(x, y, **kw):
    return LCGP(y=y, x=x, **kw), O.LCGPOracle(y=y, x=x, **kw)


def _check_loss_grad(m, o, loss_fn=None):
    f, g = m.loss_and_grad()
    fo, go = o.loss_and_grad(loss_fn)
    assert abs(f - fo) <= NLL_TOL * abs(fo), (f, fo)
    assert np.max(np.abs(g - go)) <= GRAD_TOL * np.max(np.abs(go))
    o1 = 0
    for v in m.trainable_variables:          # per parameter block, so a small block cannot hide
        k = v.numel()
        den = np.max(np.abs(go[o1:o1 + k]))
        assert np.max(np.abs(g[o1:o1 + k] - go[o1:o1 + k])) <= GRAD_TOL * max(den, 1e-3 * np.max(np.abs(go)))
        o1 += k


@pytest.mark.parametrize('case', ['rep1d_q3', 'rep1d_q2', 'rep3d', 'rep_ragged_groups', 'rep_nostd_nonrobust',
                                  'rep_n129', 'rep_n700_d5', 'rep_q1'])
def test_rep_objective_gradient(case):
    if case.startswith('rep1d'):
        x, y, _, _ = synthetic.rep1d_skewed(); kw = dict(q=int(case[-1]), submethod='rep')
    elif case == 'rep3d':
        x, y, _ = synthetic.rep3d(); kw = dict(q=3, submethod='rep')
    elif case == 'rep_ragged_groups':
        x, y, _ = make_ragged_rep_data(seed=2, n_unique=90, p=6, d=3); kw = dict(q=4, submethod='rep', diag_error_structure=[2, 1, 3])
    elif case == 'rep_nostd_nonrobust':
        x, y, _ = make_ragged_rep_data(seed=3, n_unique=60, p=4, d=2)
        kw = dict(q=2, submethod='rep', rep_standardize_ybar=False, robust_mean=False)
    elif case == 'rep_n129':
        x, y, _ = make_ragged_rep_data(seed=4, n_unique=129, p=5, d=4); kw = dict(q=3, submethod='rep')
    elif case == 'rep_n700_d5':
        x, y, _ = make_ragged_rep_data(seed=5, n_unique=700, p=12, d=5); kw = dict(q=4, submethod='rep')
    else:
        x, y, _ = make_ragged_rep_data(seed=6, n_unique=50, p=3, d=1); kw = dict(q=1, submethod='rep')
    m, o = _pair(x, y, **kw)
    _check_loss_grad(m, o)                    # init_params point
    move_params(m, o)
    _check_loss_grad(m, o)                    # off-init point
    assert float(m.neglpost_rep()) == pytest.approx(float(o.neglpost_rep().detach()), rel=NLL_TOL)


@pytest.mark.parametrize('n,p,d,q', [(50, 4, 2, 4), (150, 6, 4, 3), (300, 5, 3, 2)])
def test_full_objective_gradient(n, p, d, q):
    x, y = make_full_data(seed=n, n=n, p=p, d=d)
    m, o = _pair(x, y, q=q, submethod='full')
    _check_loss_grad(m, o, o.neglpost_chol)
    move_params(m, o)
    _check_loss_grad(m, o, o.neglpost_chol)
    # the literal eigh form (lcgp.py:650-664) agrees with the Cholesky form the kernels implement
    assert abs(float(m.neglpost()) - float(o.neglpost().detach())) <= 1e-9 * abs(float(o.neglpost().detach()))


PAD_TAILS = [1, 31, 33, 64, 96, 97, 121, 127]


def _check_padded_tail(tail):
    n = 256 + tail
    x, y = make_full_data(seed=tail, n=n, p=4, d=3)
    m, o = _pair(x, y, q=2, submethod='full')
    move_params(m, o)
    _check_loss_grad(m, o, o.neglpost_chol)
    x0 = np.random.default_rng(tail).uniform(0, 1, (70, 3))
    for a, b in zip(m.predict(x0), o.predict(torch.as_tensor(x0))):
        assert rel(a, b) <= PRED_TOL


@pytest.mark.gpu
@pytest.mark.parametrize('tail', PAD_TAILS)
def test_every_padded_tail_of_the_last_block(tail):
    """n = 2 * 128 + tail: the kernels skip the MMA steps that only multiply the identity pad of the last 128-block (K steps
    beyond n, pad rows of the last block row, pad columns of the last block column, csrc/gemm_dmma.cuh StepMask) -- one
    case per residue class the masks distinguish (8-row groups, 32-wide K steps / column chunks): objective + gradient +
    prediction against the oracle (persistent Cholesky + inverse kernel)."""
    _check_padded_tail(tail)


def _padded_tails_in_subprocess(**env_over):
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ('import sys; sys.path[:0] = [%r, %r]\n'
            'import test_gpu_parity as T\n'
            'for t in T.PAD_TAILS: T._check_padded_tail(t)\n'
            'print("PAD_TAILS_OK")\n') % (root, os.path.join(root, 'tests'))
    r = subprocess.run([sys.executable, '-c', code], env=dict(os.environ, **env_over), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and 'PAD_TAILS_OK' in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.gpu
def test_every_padded_tail_through_the_launch_chain():
    """The same cases with LCGP_POTRF=panels (SyrkJob / TrsmJob / TrtriG1 / TrtriG2 masks).  The switch is read once per
    process, hence a subprocess."""
    _padded_tails_in_subprocess(LCGP_POTRF='panels')


@pytest.mark.gpu
def test_fallback_staging_engine_parity():
    """LCGP_GEMM=cpasync: the cp.async staging engine (taken automatically when the driver does not export
    cuTensorMapEncodeTiled) runs the same Jobs and epilogues without TMA, masks or the persistent kernel."""
    _padded_tails_in_subprocess(LCGP_GEMM='cpasync')


def test_committed_fixtures():
    """Known answers committed under tests/golden (independent of running the oracle on this box)."""
    for case in FIX:
        if case['name'].startswith('rep1d'):
            x, y, xt, _ = synthetic.rep1d_skewed(); x0 = xt[::40]
        elif case['name'] == 'rep3d':
            x, y, x0 = synthetic.rep3d(); x0 = x0[:8]
        else:
            x, y, x0, _ = synthetic.latent_mixture(n=150, d=4, p=6, q_true=3, seed=5, rep_choices=None, n0=8)
        m = LCGP(y=y, x=x, **case['model'])
        f, g = m.loss_and_grad()
        assert abs(f - case['init']['loss']) <= NLL_TOL * abs(f)    # 'full' fixtures: Cholesky form (make_oracle_fixtures.py)
        assert rel(g, case['init']['grad']) < GRAD_TOL
        mv = case['moved']
        m.lLmb.assign(mv['lLmb']); m.lLmb0.assign(mv['lLmb0']); m.lsigma2s.assign(mv['lsigma2s']); m.lnugGPs.assign(mv['lnugGPs'])
        f, g = m.loss_and_grad()
        assert abs(f - mv['loss']) <= NLL_TOL * abs(f) and rel(g, mv['grad']) < GRAD_TOL
        yp, ypv, ycv = m.predict(x0)
        assert rel(yp, mv['ypred']) < PRED_TOL and rel(ypv, mv['ypredvar']) < PRED_TOL and rel(ycv, mv['yconfvar']) < PRED_TOL


def test_gradient_matches_finite_differences():
    x, y, _ = make_ragged_rep_data(seed=7, n_unique=200, p=5, d=3)
    m = LCGP(y=y, x=x, q=3, submethod='rep')
    f0, g = m.loss_and_grad()
    v0 = m._flat_get()
    rng = np.random.default_rng(0)
    for _ in range(3):
        dirn = rng.standard_normal(v0.size); dirn /= np.linalg.norm(dirn)
        h = 1e-6
        m._flat_set(v0 + h * dirn); fp = float(m.loss())
        m._flat_set(v0 - h * dirn); fm = float(m.loss())
        assert abs((fp - fm) / (2 * h) - g @ dirn) <= 2e-6 * max(1.0, abs(g @ dirn))


# ---------------------------------------------------------------- a5-a8: aux quantities and prediction
@pytest.mark.parametrize('sub', ['rep', 'full'])
def test_predict_and_aux(sub):
    if sub == 'rep':
        x, y, xu = make_ragged_rep_data(seed=8, n_unique=140, p=5, d=3); xt = xu
    else:
        x, y = make_full_data(seed=8, n=140, p=5, d=3); xt = x
    m, o = _pair(x, y, q=3, submethod=sub)
    move_params(m, o)
    x0 = np.random.default_rng(1).uniform(0, 1, (37, 3))
    res, reso = m.predict(x0), o.predict(torch.as_tensor(x0))
    for a, b in zip(res, reso):
        assert a.shape == (5, 37) and rel(a, b) < PRED_TOL
    assert rel(m.ghat, o.ghat) < PRED_TOL and rel(m.gvar, o.gvar) < PRED_TOL
    assert rel(m.CinvMs, o.CinvMs) < PRED_TOL
    if sub == 'rep':
        assert rel(m.mks, o.mks) < PRED_TOL
        assert rel(m.Tks[1], o.Tks[1]) < PRED_TOL            # oracle Tks in the stable form (helpers/oracle docstring)
        assert rel(m.psi_c, o.psi_c) < 1e-12
    else:
        Th = m.Ths[0]
        assert rel(Th @ Th.T, o.Ths[0] @ o.Ths[0].T) < PRED_TOL   # Th Th^T = (C + I/d)^-1, lcgp.py:709-715
    # prediction AT the training inputs exercises the nugget-on-equal-inputs branch (covmat.py:46-53)
    res, reso = m.predict(xt), o.predict(torch.as_tensor(xt))
    for a, b in zip(res, reso):
        assert rel(a, b) < PRED_TOL
    # chunked prediction (n0 > chunk) gives the same answer
    big = np.random.default_rng(2).uniform(0, 1, (300, 3))
    g1, v1, _ = m._predict_latents(torch.as_tensor(big), m.x_unique_s if sub == 'rep' else m.x, chunk=128)
    g2, v2, _ = m._predict_latents(torch.as_tensor(big), m.x_unique_s if sub == 'rep' else m.x, chunk=4096)
    assert torch.equal(g1, g2) and torch.equal(v1, v2)


def test_fullcov_and_rep_fullcov_none():       # test_coverage_gaps.py:169-232
    x, y = make_full_data(seed=0, n=40, p=3, d=2)
    m, o = _pair(x, y, submethod='full')
    x0 = np.random.default_rng(11).uniform(0, 1, (8, 2))
    yp, ypv, ycv, full = m.predict(x0, return_fullcov=True)
    assert full.shape == (8, 3, 3)
    np.testing.assert_allclose(np.diagonal(full.numpy(), axis1=1, axis2=2).T, ypv.numpy(), rtol=1e-5, atol=1e-6)
    assert rel(full, o.predict(torch.as_tensor(x0), return_fullcov=True)[3]) < PRED_TOL
    xr, yr, _ = make_rep_data(n_unique=20, p=4, d=2, reps=3)
    mr = LCGP(y=yr, x=xr, submethod='rep', rep_standardize_ybar=False)
    out = mr.predict(x0, return_fullcov=True)
    assert len(out) == 4 and out[3] is None and torch.isfinite(out[0]).all()


@pytest.mark.parametrize('q,p,n0', [(3, 3, 8), (5, 130, 37), (32, 300, 9), (70, 257, 5), (1, 1, 4), (130, 130, 3)])
def test_fullcov_kernel_matches_formula(q, p, n0):
    """lcgp_predict_fullcov (csrc/fullcov.cu) against lcgp.py:850-857 restated in torch on the CPU: odd / even p
    (scalar and 16-byte stores), several 128-tiles, q below / above one 64-latent shared-memory chunk."""
    from lcgp_b200.model import _fullcov_cuda
    g = torch.Generator().manual_seed(q * 1000 + p)
    psi = torch.randn(q, p, dtype=torch.float64, generator=g)
    gvar = torch.rand(q, n0, dtype=torch.float64, generator=g) + 0.01
    sig2 = torch.rand(p, dtype=torch.float64, generator=g) + 0.1
    sv = torch.rand(p, dtype=torch.float64, generator=g) + 0.5
    full = _fullcov_cuda(psi, gvar, sig2, sv)
    CH = torch.einsum('kn,kp->npk', torch.sqrt(gvar), psi)
    ref = (CH @ CH.transpose(1, 2) + torch.diag(sig2)[None]) * (sv[:, None] * sv[None, :])[None]
    assert full.shape == (n0, p, p)
    assert rel(full, ref) < 1e-13
    assert torch.equal(full, full.transpose(1, 2))          # exactly symmetric: same products in the same order
    small = _fullcov_cuda(psi, gvar, sig2, sv, chunk_bytes=8 * p * p * 2)      # chunked over test points
    assert torch.equal(small, full)


# ---------------------------------------------------------------- evaluation plans (CUDA graphs)
def test_plan_replays_a_graph_and_equals_the_launch_by_launch_path():
    """Small problems evaluate through lcgp_plan_run: the capture must succeed (a real graph, not the eager
    fallback) and replaying it at new parameter values must give the bits of the launch-by-launch call."""
    x, y, _ = make_ragged_rep_data(seed=31, n_unique=150, p=5, d=3)
    mg = LCGP(y=y, x=x, q=4, submethod='rep')
    me = LCGP(y=y, x=x, q=4, submethod='rep')
    assert mg.engine.use_plans
    me.engine.use_plans = False
    from lcgp_b200 import _cabi
    for step in range(4):
        if step:
            move_params(mg, seed=100 + step); move_params(me, seed=100 + step)
        c0 = int(_cabi.lib().lcgp_launch_count())
        fg, gg = mg.loss_and_grad()
        c1 = int(_cabi.lib().lcgp_launch_count())
        fe, ge = me.loss_and_grad()
        c2 = int(_cabi.lib().lcgp_launch_count())
        assert fg == fe and np.array_equal(gg, ge)
        if step:                                   # replay = one graph launch; the eager path launches every kernel
            assert c1 - c0 == 1 and c2 - c1 > 10      # (the persistent Cholesky kernel is ONE launch per stream group)
    assert mg.engine.plan_is_graph()
    assert float(mg.loss()) == float(me.loss())    # the objective-only plan (flags without the gradient bit)
    x0 = np.random.default_rng(3).uniform(0, 1, (9, 3))
    for a, b in zip(mg.predict(x0), me.predict(x0)):   # prediction reads the factor the replay left in the workspace
        assert torch.equal(a, b)


def test_batched_emulators_on_threads_match_sequential_fits():
    """fit_emulators (BASELINE config 5 path: several host threads, one stream and one graph plan each) gives
    exactly the parameters of one-at-a-time fits."""
    from lcgp_b200 import fit_emulators
    data = [make_full_data(seed=40 + i, n=140, p=4, d=2) for i in range(5)]
    mk = dict(q=2, submethod='full')
    res = fit_emulators(data, mk, fit_options=dict(maxiter=6), threads_per_gpu=3, engine='threads')
    for (x, y), r in zip(data, res):
        m = LCGP(y=y, x=x, **mk)
        m.fit(maxiter=6)
        assert np.array_equal(m.get_param()[0].numpy(), r['lLmb']) and m.n_evals + 0 >= 1
        assert abs(float(m.loss()) - r['loss']) == 0.0


def test_batched_engine_equals_one_emulator_at_a_time():
    """lcgp_problem.n_emu: E emulators (different data, identical shapes) evaluated by ONE call give, block by block,
    the `out` vectors of E separate calls; n spans several 128-blocks, q is not a multiple of 8, replicated data with
    per-emulator 1/n scale and sum log r."""
    from lcgp_b200.batched import BatchedEngine
    E = 3
    models = []
    for e in range(E):
        x, y, _ = make_ragged_rep_data(seed=60 + e, n_unique=300, p=6, d=3)
        models.append(LCGP(y=y, x=x, q=3, submethod='rep'))
    # identical shapes are required: same number of unique inputs
    assert len({int(m.n) for m in models}) == 1
    for e, m in enumerate(models):
        move_params(m, seed=70 + e)
    eng = BatchedEngine(models)
    pars = [m.get_param() for m in models]
    st = lambda k: torch.stack([p[k] for p in pars])
    out = eng.evaluate(st(0), st(1), st(3), st(2), True).clone()
    for e, m in enumerate(models):
        lLmb, lLmb0, lsig_p, lnug = pars[e]
        one = m.engine.evaluate(lLmb, lLmb0, lnug, lsig_p, True)
        assert rel(out[e], one) < 1e-14, (e, rel(out[e], one))
        assert abs(float(out[e, 0]) - float(one[0])) <= 1e-15 * abs(float(one[0]))


def test_fit_emulators_lockstep_equals_sequential_fits():
    """engine='lockstep': every emulator's SciPy L-BFGS-B state machine is served by batched evaluations; the fitted
    parameters, evaluation counts and final objectives are those of one-at-a-time LCGP.fit() calls (same routine,
    same objective bits), including after half of the batch has converged and the batch is compacted."""
    from lcgp_b200 import fit_emulators
    data = [make_full_data(seed=80 + i, n=140, p=4, d=2) for i in range(5)]
    mk = dict(q=2, submethod='full')
    opts = [dict(maxiter=6), dict(maxiter=40)]
    for o in opts:
        res = fit_emulators(data, mk, fit_options=o, engine='lockstep')
        for (x, y), r in zip(data, res):
            m = LCGP(y=y, x=x, **mk)
            m.fit(**o)
            assert m.n_evals == r['n_evals'] and m.opt_result.nit == r['nit']
            # (the batched and the one-emulator evaluations agree to ~1e-15, not bit for bit; L-BFGS amplifies that by
            # roughly 10^3 per 20 iterations, see test_fit_matches_oracle_under_shared_optimizer)
            tol = 1e-11 if o['maxiter'] <= 6 else 1e-7
            assert rel(m.get_param()[0].numpy(), r['lLmb']) < tol and rel(m.lsigma2s.numpy(), r['lsigma2s']) < tol
            assert abs(float(m.loss()) - r['loss']) <= 1e-9 * abs(r['loss'])


def test_sharded_evaluation_with_ranks_emulated_on_one_gpu():
    """The multi-rank path without NCCL (one GPU is enough): W 'ranks' are W engines over the round-robin latent shards
    of one model, each result packed by lcgp_pack_sharded; the SUM of the packed vectors (what the all-reduce computes)
    must equal the unsharded evaluation and the oracle; a rank's failed latent shows up in the last slot."""
    from lcgp_b200.model import CudaEngine
    x, y, _ = make_ragged_rep_data(seed=14, n_unique=260, p=7, d=3)
    m, o = _pair(x, y, q=5, submethod='rep', diag_error_structure=[3, 4])
    move_params(m, o)
    lLmb, lLmb0, lsig_p, lnug = m.get_param()
    q, d, p = 5, 3, 7
    ref = m.engine.evaluate(lLmb, lLmb0, lnug, lsig_p, True)[:1 + p + q * d + 2 * q]
    data = m._problem_data()
    for W in (2, 3):
        total = torch.zeros(1 + p + q * d + 2 * q + 1, dtype=DT, device='cuda')
        for rank in range(W):
            idx = torch.arange(q)[rank::W]
            eng = CudaEngine(phi_loc=m.phi[:, idx], D_loc=m.diag_D[idx], include_host_terms=(rank == 0), **data)
            eng.evaluate_device(lLmb[idx], lLmb0[idx], lnug[idx], lsig_p, True)
            loc = torch.full((q,), -1, dtype=torch.int32)
            loc[idx] = torch.arange(idx.numel(), dtype=torch.int32)
            flat = torch.full_like(total, float('nan'))
            eng.pack_sharded(loc.cuda(), q, flat)
            total += flat
        assert float(total[-1]) == 0.0
        assert rel(total[:-1].cpu(), ref) < 1e-13
    # against the oracle, through the same chain rule the model applies
    fo, go = o.loss_and_grad()
    assert abs(float(ref[0]) - fo) <= NLL_TOL * abs(fo)
    # a NaN parameter on one 'rank' is counted in the last slot
    bad = lLmb0.clone(); bad[1] = float('nan')
    idx = torch.arange(q)[1::2]
    eng = CudaEngine(phi_loc=m.phi[:, idx], D_loc=m.diag_D[idx], include_host_terms=False, **data)
    eng.evaluate_device(lLmb[idx], bad[idx], lnug[idx], lsig_p, True)
    loc = torch.full((q,), -1, dtype=torch.int32); loc[idx] = torch.arange(idx.numel(), dtype=torch.int32)
    flat = torch.zeros(1 + p + q * d + 2 * q + 1, dtype=DT, device='cuda')
    eng.pack_sharded(loc.cuda(), q, flat)
    assert float(flat[-1]) == 1.0


def test_predict_outputs_kernel_matches_the_formulas():
    """lcgp_predict_outputs against lcgp.py:915-926 restated in torch: ragged p / n0 / q, with and without scale / shift."""
    L = _cabi.lib()
    g = torch.Generator().manual_seed(9)
    for p, q, n0, std in [(7, 3, 5, True), (130, 33, 300, True), (2, 1, 129, False), (64, 64, 1, True)]:
        Psi = torch.randn(p, q, dtype=DT, generator=g); gh = torch.randn(q, n0, dtype=DT, generator=g)
        gv = torch.rand(q, n0, dtype=DT, generator=g); nv = torch.rand(p, dtype=DT, generator=g)
        sc = torch.rand(p, dtype=DT, generator=g) + 0.5 if std else None
        sh = torch.randn(p, dtype=DT, generator=g) if std else None
        c = lambda t: None if t is None else t.cuda().contiguous()
        Pd, ghd, gvd, nvd, scd, shd = map(c, (Psi, gh, gv, nv, sc, sh))
        outs = [torch.full((p, n0), float('nan'), dtype=DT, device='cuda') for _ in range(3)]
        ptr = lambda t: None if t is None else t.data_ptr()
        assert L.lcgp_predict_outputs(Pd.data_ptr(), ghd.data_ptr(), gvd.data_ptr(), nvd.data_ptr(), ptr(scd), ptr(shd), p, q, n0,
                                      outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(), _cabi.stream_ptr()) == 0
        s1 = torch.ones(p, dtype=DT) if sc is None else sc
        s0 = torch.zeros(p, dtype=DT) if sh is None else sh
        mean, conf = Psi @ gh, (Psi ** 2) @ gv
        assert rel(outs[0].cpu(), mean * s1[:, None] + s0[:, None]) < 1e-13
        assert rel(outs[2].cpu(), conf * s1[:, None] ** 2) < 1e-13
        assert rel(outs[1].cpu(), (conf + nv[:, None]) * s1[:, None] ** 2) < 1e-13


# ---------------------------------------------------------------- f-1: preprocessing on the device
@pytest.mark.parametrize('sub,robust', [('rep', True), ('rep', False), ('full', True), ('full', False)])
def test_device_preprocessing_equals_host_preprocessing(sub, robust):
    """csrc/prep.cu (replicate means, nearest-rank median / MAD, standardisation) vs the host pipeline of
    lcgp.py:312-324, 358-395: identical bits for every attribute; the objective's row sums w to 1e-14."""
    if sub == 'rep':
        x, y, _ = make_ragged_rep_data(seed=21, n_unique=333, p=7, d=3)
    else:
        x, y = make_full_data(seed=22, n=301, p=6, d=2)
    md = LCGP(y=y, x=x, q=3, submethod=sub, robust_mean=robust, device_preprocess=True)
    mh = LCGP(y=y, x=x, q=3, submethod=sub, robust_mean=robust, device_preprocess=False)
    names = ('ybar', 'ybar_mean', 'ybar_std', 'ybar_s') if sub == 'rep' else ('y', 'ymean', 'ystd')
    for nm in names:
        assert torch.equal(getattr(md, nm), getattr(mh, nm)), nm
    assert torch.equal(md.phi, mh.phi) and torch.equal(md.diag_D, mh.diag_D)
    dd, dh = md._problem_data(), mh._problem_data()
    assert dd['YR'].is_cuda and torch.equal(dd['YR'].cpu(), dh['YR'])
    assert rel(dd['w'].cpu(), dh['w']) < 1e-14
    fd, gd = md.loss_and_grad()
    fh, gh = mh.loss_and_grad()
    assert abs(fd - fh) <= 1e-13 * abs(fh) and np.max(np.abs(gd - gh)) <= 1e-12 * np.max(np.abs(gh))


def test_gram_basis_equals_svd_basis():
    """Constructor at large p (f-1): phi from the eigen-decomposition of the p x p Gram matrix (device product + host
    eigh) against phi from the p x n SVD the reference computes (lcgp.py:464-479), columns sign-fixed (SURVEY B-15);
    diag_D, the objective and its gradient agree to the north-star tolerances."""
    x, y, x0, _ = synthetic.latent_mixture(n=900, d=4, p=320, q_true=8, seed=77, rep_choices=(1, 2), n0=16)
    mg = LCGP(y=y, x=x, q=8, submethod='rep', device_preprocess=True)     # p >= 256, 4 q <= p <= n: Gram path
    ms = LCGP(y=y, x=x, q=8, submethod='rep', device_preprocess=False)    # host SVD
    calls = []
    orig = mg._gram_basis
    mg._gram_basis = lambda Y, q: (calls.append(1), orig(Y, q))[1]
    mg.init_phi()
    assert calls, 'the Gram path was not taken'
    sg = torch.sign(mg.phi[mg.phi.abs().argmax(dim=0), torch.arange(8)])
    ss = torch.sign(ms.phi[ms.phi.abs().argmax(dim=0), torch.arange(8)])
    assert rel(mg.phi * sg, ms.phi * ss) < 1e-10
    assert rel(mg.diag_D, ms.diag_D) < 1e-11
    fg, gg = mg.loss_and_grad()
    fs, gs = ms.loss_and_grad()
    assert abs(fg - fs) <= NLL_TOL * abs(fs) and np.max(np.abs(gg - gs)) <= GRAD_TOL * np.max(np.abs(gs))
    # ill-separated request (all p components of a rank-deficient matrix): the Gram path declines
    Yb = torch.randn(300, 5, dtype=DT) @ torch.randn(5, 400, dtype=DT)
    assert mg._gram_basis(Yb, 20) == (None, None)


@pytest.mark.parametrize('m', [1, 2, 7, 256, 1001, 8000])
def test_row_select_is_the_nearest_rank_percentile(m):
    """lcgp_prep_row_select == sorted row at index round((m-1)/2) (tfp 'nearest' percentile), incl. ties, negative
    values, zeros and the |y - c| form."""
    from lcgp_b200 import _cabi
    g = torch.Generator().manual_seed(m)
    Y = torch.randn(5, m, dtype=torch.float64, generator=g)
    Y[1] = torch.round(Y[1] * 2) / 2                 # many ties, exact zeros
    Y[2] = -torch.abs(Y[2])                          # all negative
    Y[3] = Y[3] * 1e-300                             # tiny magnitudes
    Y[4] = 3.25
    k = int(np.round((m - 1) * 0.5))
    Yd = Y.cuda()
    out = torch.empty(5, dtype=torch.float64, device='cuda')
    _cabi.check(_cabi.lib().lcgp_prep_row_select(Yd.data_ptr(), None, 5, m, k, out.data_ptr(), _cabi.stream_ptr()), 'select')
    want = torch.sort(Y, dim=1).values[:, k]
    assert torch.equal(out.cpu(), want)
    c = want.cuda()
    _cabi.check(_cabi.lib().lcgp_prep_row_select(Yd.data_ptr(), c.data_ptr(), 5, m, k, out.data_ptr(), _cabi.stream_ptr()), 'select')
    assert torch.equal(out.cpu(), torch.sort(torch.abs(Y - want[:, None]), dim=1).values[:, k])


@pytest.mark.parametrize('sub', ['rep', 'full'])
def test_gradient_wrt_latent_basis_matches_oracle_autograd(sub):
    """lcgp_grad_phi (no reference counterpart) against autograd through the oracle's objective with
    phi a leaf and diag_D = sum_j phi_jk^2 a function of it; n spans several 128-blocks with a ragged last one."""
    if sub == 'rep':
        x, y, _ = make_ragged_rep_data(seed=17, n_unique=300, p=6, d=3)
    else:
        x, y = make_full_data(seed=18, n=290, p=5, d=2)
    m, o = _pair(x, y, q=3, submethod=sub)
    move_params(m, o, seed=5)
    g = m.grad_phi()
    phi = o.phi.clone().requires_grad_(True)
    o.phi, o.diag_D = phi, (phi ** 2).sum(dim=0)
    o.loss().backward()
    assert g.shape == phi.shape
    assert rel(g, phi.grad) < GRAD_TOL, rel(g, phi.grad)


def test_auxiliary_kernels_respect_their_output_bounds():
    """Guard bands around every output of lcgp_predict_fullcov and the lcgp_prep_* kernels (ragged sizes: p, n0,
    n not multiples of any tile) stay untouched (compute-sanitizer is not available on this pool)."""
    L = _cabi.lib()
    dev, G, SENT = torch.device('cuda'), 257, -12345.678
    st = _cabi.stream_ptr()

    def guarded(numel):
        buf = torch.full((numel + 2 * G,), SENT, dtype=torch.float64, device=dev)
        return buf, buf[G:G + numel]

    def intact(buf, numel):
        return bool((buf[:G] == SENT).all()) and bool((buf[G + numel:] == SENT).all())

    g = torch.Generator(device='cuda').manual_seed(5)
    for q, p, n0 in [(5, 131, 19), (3, 7, 33), (70, 129, 3)]:
        psi = torch.randn(q, p, dtype=torch.float64, device=dev, generator=g)
        gv = torch.rand(q, n0, dtype=torch.float64, device=dev, generator=g)
        s2 = torch.rand(p, dtype=torch.float64, device=dev, generator=g)
        sv = torch.rand(p, dtype=torch.float64, device=dev, generator=g) + 0.5
        buf, out = guarded(n0 * p * p)
        _cabi.check(L.lcgp_predict_fullcov(psi.data_ptr(), gv.data_ptr(), s2.data_ptr(), sv.data_ptr(), q, p, n0,
                                           out.data_ptr(), st), 'fullcov')
        torch.cuda.synchronize()
        assert intact(buf, n0 * p * p) and bool(torch.isfinite(out).all()) and not bool((out == SENT).any())
    p, N, n = 9, 1003, 377
    rng = np.random.default_rng(0)
    inv = np.concatenate([np.arange(n), rng.integers(0, n, N - n)]); rng.shuffle(inv)
    order = torch.as_tensor(np.argsort(inv, kind='stable').astype(np.int32)).to(dev)
    off = torch.as_tensor(np.concatenate([[0], np.cumsum(np.bincount(inv, minlength=n))]).astype(np.int32)).to(dev)
    y = torch.randn(p, N, dtype=torch.float64, device=dev, generator=g)
    bybar, ybar = guarded(p * n)
    _cabi.check(L.lcgp_prep_segment_mean(y.data_ptr(), order.data_ptr(), off.data_ptr(), p, N, n, ybar.data_ptr(), st), 'segmean')
    bc, c = guarded(p)
    _cabi.check(L.lcgp_prep_row_select(ybar.data_ptr(), None, p, n, n // 2, c.data_ptr(), st), 'select')
    bs, sp = guarded(p)
    _cabi.check(L.lcgp_prep_row_select(ybar.data_ptr(), c.data_ptr(), p, n, n // 2, sp.data_ptr(), st), 'select')
    r = torch.rand(n, dtype=torch.float64, device=dev, generator=g) + 1.0
    bys, ys = guarded(p * n); byr, yr = guarded(p * n); bw, w = guarded(p)
    _cabi.check(L.lcgp_prep_standardize(ybar.data_ptr(), c.data_ptr(), sp.data_ptr(), r.data_ptr(), p, n, ys.data_ptr(),
                                        yr.data_ptr(), w.data_ptr(), st), 'standardize')
    torch.cuda.synchronize()
    for b, m in ((bybar, p * n), (bc, p), (bs, p), (bys, p * n), (byr, p * n), (bw, p)):
        assert intact(b, m)
    Yb = ybar.view(p, n)
    assert torch.equal(c, torch.sort(Yb, dim=1).values[:, n // 2])
    assert torch.allclose(w, ((Yb - c[:, None]) / sp[:, None]) ** 2 @ r, rtol=1e-13)


# ---------------------------------------------------------------- a10: fit
@pytest.mark.parametrize('optimizer', ['L-BFGS-B', 'torch-lbfgs'])
def test_fit_matches_oracle_under_shared_optimizer(optimizer):
    """Oracle and CUDA path under ONE optimizer, start and iteration budget (SURVEY 7 'hard parts').
    L-BFGS amplifies the ~1e-16 differences between the two implementations by roughly 10^3 per 20
    iterations on these weakly identified objectives (tools/fit_parity_probe.py: 1e-12 after 10
    iterations, 1e-9 after 30, divergent trajectories after several hundred), so the north-star 1e-6 on
    fitted hyper-parameters is asserted for a bounded budget; parity along the whole path is covered by
    test_parity_along_fit_path, converged fits by the notebook test below."""
    x, y, xu = make_ragged_rep_data(seed=9, n_unique=80, p=4, d=2)
    m, o = _pair(x, y, q=3, submethod='rep')
    l0 = float(m.loss())
    if optimizer == 'L-BFGS-B':
        m.fit(optimizer=optimizer, maxiter=30)
        o.fit(method='L-BFGS-B', maxiter=30)
        assert m.n_evals - 1 == o.opt_result.nfev          # same number of closure calls (+1 for l0 above)
    else:
        m.fit(optimizer=optimizer, max_iter=20)
        o.fit(method='torch-lbfgs', max_iter=20)
    l1, lo1 = float(m.loss()), float(o.loss().detach())
    assert l1 <= l0 + 1e-3                                       # test_rep.py:161-171
    assert abs(l1 - lo1) <= 1e-9 * abs(lo1)
    for a, b in ((m.lLmb, o.lLmb), (m.lLmb0, o.lLmb0), (m.lsigma2s, o.lsigma2s), (m.lnugGPs, o.lnugGPs)):
        assert rel(a.numpy(), b.detach().numpy()) < PRED_TOL     # 1e-6, all four parameter blocks
    x0 = np.random.default_rng(5).uniform(0, 1, (25, 2))
    for a, b in zip(m.predict(x0), o.predict(torch.as_tensor(x0))):
        assert rel(a, b) < PRED_TOL
    for t in m.get_param():
        assert torch.isfinite(t).all()


def test_parity_along_fit_path():
    """Objective and gradient parity at the iterates of a full CUDA-path fit (not only at init)."""
    x, y, _ = synthetic.rep3d()
    m, o = _pair(x, y, q=3, submethod='rep')
    path = []
    orig = m.loss_and_grad

    def spy():
        path.append(m._flat_get())
        return orig()
    m.loss_and_grad = spy
    m.fit(maxiter=120)
    m.loss_and_grad = orig
    assert len(path) > 30
    for v in [path[i] for i in np.linspace(0, len(path) - 1, 8).astype(int)]:
        m._flat_set(v); o._flat_set(v)
        f, g = m.loss_and_grad()
        fo, go = o.loss_and_grad()
        assert abs(f - fo) <= NLL_TOL * abs(fo)
        # near the optimum the gradient is a ~1e-4 residue of cancelling O(1) terms (absolute difference
        # measured there: 5e-12): relative 1e-8, with an absolute floor at the objective's own tolerance
        assert np.max(np.abs(g - go)) <= max(GRAD_TOL * np.max(np.abs(go)), NLL_TOL * max(1.0, abs(fo)))


def test_notebook_goldens_through_cuda_path():
    """End to end on the notebook case: the reference's stored metrics, produced by the CUDA path."""
    xtr, ytr, xte, ytrue = synthetic.rep1d_skewed()
    m = LCGP(y=ytr, x=xtr, q=3, submethod='rep', diag_error_structure=[1, 1, 1], robust_mean=True)
    np.testing.assert_allclose(m.diag_D.numpy(), GOLD['diag_D'], atol=5e-9)
    assert abs(float(m.loss()) - 0.2793219611474477) < 1e-11
    m.fit()
    np.testing.assert_allclose(m.lLmb.numpy().ravel(), GOLD['fitted_lengthscales'], rtol=1e-3)
    np.testing.assert_allclose(m.lsigma2s.numpy(), GOLD['fitted_lsigma2s'], rtol=1e-3)
    # converged fit vs the oracle's converged fit (same SciPy defaults, 72 evaluations each): the
    # identifiable blocks agree to 1e-5 (measured 2e-6 / 5e-8); see test_fit_matches_oracle_... for why not tighter
    o = O.LCGPOracle(y=ytr, x=xtr, q=3, submethod='rep', diag_error_structure=[1, 1, 1], robust_mean=True)
    o.fit()
    assert rel(m.lLmb.numpy(), o.lLmb.detach().numpy()) < 1e-5 and rel(m.lsigma2s.numpy(), o.lsigma2s.detach().numpy()) < 1e-5
    assert abs(float(m.loss()) - float(o.loss().detach())) <= 1e-9 * abs(float(o.loss().detach()))
    yp, ypv, ycv = (t.numpy() for t in m.predict(xte))
    assert round(float(evaluation.rmse(ytrue, yp)), 4) == GOLD['rmse']
    assert round(float(evaluation.normalized_rmse(ytrue, yp)), 4) == GOLD['nrmse']
    cov, width = evaluation.intervalstats(ytrue, yp, ycv)
    assert round(float(cov), 3) == GOLD['coverage'] and round(float(width), 4) == GOLD['width']
    assert abs(evaluation.dss(ytrue, yp, ycv, use_diag=True) - GOLD['dss']) < 2e-4
    assert np.all(ypv > 0) and np.all(ycv <= ypv + 1e-9)         # test_rep.py:203-232


def test_example_harness_trains_and_predicts():
    """LCGPRun.train / predict (docs/call_model.py:59-86) through the CUDA path: numpy outputs, (p, n0) or transposed
    with as_pxn, 4-tuple with the full covariance in 'full' mode, prediction at the training inputs."""
    from lcgp_b200 import LCGPRun
    x, y = make_full_data(seed=5, n=60, p=3, d=2)
    run = LCGPRun(runno='t', data=dict(xtrain=x, ytrain=y, xtest=x[:7] + 0.01, ytest=y[:, :7]), submethod='full', num_latent=2)
    run.define_model()
    run.train()
    ym, ypv, ycv = run.predict()
    assert isinstance(ym, np.ndarray) and ym.shape == (3, 7) and (ypv > 0).all() and (ycv <= ypv + 1e-9).all()
    t = run.predict(as_pxn=True)
    assert t[0].shape == (7, 3) and np.array_equal(t[0].T, ym)
    out = run.predict(train=True, return_fullcov=True)
    assert len(out) == 4 and out[0].shape == (3, 60) and out[3].shape == (60, 3, 3)


def test_reference_behaviour_contract():
    """Shape / invariants the reference tests assert after fit+predict (test_training.py, test_rep.py)."""
    x, y = make_full_data(seed=1, n=50, p=4, d=2)
    m = LCGP(y=y, x=x, submethod='full')
    m.fit()
    yp, ypv, ycv = m.predict(x0=x)                              # at the training inputs, test_training.py:15
    assert yp.shape == (4, 50) and torch.isfinite(yp).all() and (ypv > 0).all() and (ycv <= ypv + 1e-9).all()
    assert len(m.get_param()) == 4
    xr, yr, xu = make_rep_data(seed=0, n_unique=20, p=4, d=2, reps=3)
    mr = LCGP(y=yr, x=xr, submethod='rep')
    mr.submethod = 'bogus'
    with pytest.raises(KeyError):
        mr.predict(x0=mr.x_unique)                              # test_coverage_gaps.py:138-147
    with pytest.raises(ValueError):
        mr.loss()
    mr.submethod = 'rep'
    mr.CinvMs = torch.full((mr.q, int(mr.n)), float('nan'), dtype=DT); mr.Tks = None
    mr.compute_aux_predictive_quantities()                      # test_coverage_gaps.py:149-166
    assert mr.Tks is not None and not torch.isnan(mr.CinvMs).any()
    mr.fit()                                                    # a second predict after fit must not reuse stale aux
    a = mr.predict(xu)[0]
    mr.lLmb.assign(mr.lLmb.numpy() * 1.3)
    assert rel(mr.predict(xu)[0], a) > 1e-6


# ---------------------------------------------------------------- oracle parity at BASELINE's named shapes
@pytest.mark.parametrize('cfg', ['cfg3_rep', 'cfg3_full', 'cfg5_one'])
def test_oracle_parity_at_named_config_shapes(cfg):
    """BASELINE.json configs 3 (n=2000, d=8, p=500, q=10; 'rep' and 'full') and 5 (one emulator: n=1024, d=6,
    p=64, q=8, 'full') at their NAMED sizes against the CPU oracle (lcgp.py:554-666, 808-930): objective 1e-10,
    gradient 1e-8 (per parameter block), predictions 1e-6, at the init_params point and at one moved point."""
    torch.set_num_threads(os.cpu_count() or 1)
    x, y, x0, _, mk = synthetic.make_config(cfg)
    m = LCGP(y=y, x=x, **mk)
    o = O.LCGPOracle(y=y, x=x, skip_xnorm=True, **mk)
    want = dict(cfg3_rep=(2000, 8, 500, 10), cfg3_full=(2000, 8, 500, 10), cfg5_one=(1024, 6, 64, 8))[cfg]
    assert (int(m.n), int(m.d), int(m.p), int(m.q)) == want
    fn = o.neglpost_chol if mk['submethod'] == 'full' else None
    _check_loss_grad(m, o, fn)                # init_params point
    move_params(m, o)
    _check_loss_grad(m, o, fn)                # moved point
    x0 = x0[:32]
    for a, b in zip(m.predict(x0), o.predict(torch.as_tensor(x0))):
        assert a.shape == (want[2], 32) and rel(a, b) < PRED_TOL


# ---------------------------------------------------------------- size-independent properties at scale
def test_properties_at_config3_scale():
    """n=2000, d=8 (BASELINE config 3 shape, fewer latents): properties that need no CPU oracle run --
    L L^T = A, U L^T = I on block rows, gradient vs central differences, predvar bounds."""
    L = _cabi.lib(); dev = torch.device('cuda')
    n, d, q = 2000, 8, 2
    rng = np.random.default_rng(2000)
    X = torch.as_tensor(rng.uniform(0, 1, (n, d)), device=dev); sr = torch.sqrt(torch.as_tensor(rng.integers(1, 5, n).astype(float), device=dev))
    ell = torch.full((q, d), 0.8, dtype=DT, device=dev); s0 = torch.ones(q, dtype=DT, device=dev)
    nug = torch.full((q,), float(np.exp(-10.0)), dtype=DT, device=dev); D = torch.tensor([0.5, 2.0], dtype=DT, device=dev)
    npad = _cabi.padded(n); nb = npad // 128; st = _cabi.stream_ptr()
    F = torch.empty((q, npad, npad), dtype=DT, device=dev)
    L.lcgp_build_A(X.data_ptr(), sr.data_ptr(), n, d, ell.data_ptr(), s0.data_ptr(), nug.data_ptr(), D.data_ptr(), q, F.data_ptr(), npad, st)
    A = torch.tril(F) + torch.tril(F, -1).transpose(1, 2)
    DL = torch.zeros((q, nb, 128, 128), dtype=DT, device=dev); DU = torch.zeros_like(DL)
    info = torch.zeros(q, dtype=torch.int32, device=dev)
    assert _potrf(L, F, npad, q, DL, DU, None, info, st) == 0
    Lf = torch.tril(F)
    resid = (Lf @ Lf.transpose(1, 2) - A).abs().max() / A.abs().max()
    assert float(resid) < 1e-13 and info.cpu().tolist() == [0, 0]
    sb = int(L.lcgp_trtri_scratch_bytes(npad, q)); scr = torch.empty(sb // 8, dtype=DT, device=dev)
    L.lcgp_trtri_batched(F.data_ptr(), npad, q, DL.data_ptr(), DU.data_ptr(), scr.data_ptr(), sb, st)
    U = torch.triu(F, 1)
    for I in range(nb):
        U[:, I * 128:(I + 1) * 128, I * 128:(I + 1) * 128] = DU[:, I]
    eye = torch.eye(npad, dtype=DT, device=dev)
    assert float((U @ Lf.transpose(1, 2) - eye).abs().max()) < 1e-10          # U = L^-T
    # model-level: gradient vs central differences and variance bounds at n = 2000
    x, y, x0, _, mk = synthetic.make_config('cfg3_rep', p=40)
    m = LCGP(y=y, x=x, q=3, submethod='rep')
    f0, g = m.loss_and_grad()
    v0 = m._flat_get()
    dirn = np.random.default_rng(1).standard_normal(v0.size); dirn /= np.linalg.norm(dirn)
    h = 1e-5
    m._flat_set(v0 + h * dirn); fp = float(m.loss())
    m._flat_set(v0 - h * dirn); fm = float(m.loss())
    m._flat_set(v0)
    assert abs((fp - fm) / (2 * h) - g @ dirn) <= 1e-5 * max(1.0, abs(g @ dirn))
    yp, ypv, ycv = m.predict(x0)
    assert torch.isfinite(yp).all() and (ypv > 0).all() and (ycv <= ypv + 1e-9).all() and (ycv > -1e-9).all()


# ---------------------------------------------------------------- config 5: batched independent fits
def test_fit_emulators_matches_sequential_fits():
    from lcgp_b200 import fit_emulators, perturbed_restart
    data = [make_full_data(seed=100 + i, n=90, p=4, d=2) for i in range(6)]
    mk = dict(q=2, submethod='full')
    opts = dict(maxiter=15)
    res = fit_emulators(data, mk, fit_options=opts, threads_per_gpu=3)
    assert [r['index'] for r in res] == list(range(6))
    for i in (0, 5):                                   # concurrency must not change any emulator's answer
        m = LCGP(y=data[i][1], x=data[i][0], **mk)
        m.fit(**opts)
        assert abs(float(m.loss()) - res[i]['loss']) <= 1e-9 * abs(res[i]['loss'])
        assert rel(m.lLmb.numpy(), res[i]['lLmb']) < 1e-7
    # multi-start on one data set: different starts, all finite
    res2, models = fit_emulators([data[0]] * 3, mk, fit_options=opts, threads_per_gpu=3,
                                 init_hooks=[None, perturbed_restart(1), perturbed_restart(2)], return_models=True)
    assert len(models) == 3 and all(np.isfinite(r['loss']) for r in res2)


# ---------------------------------------------------------------- edge cases of the C-ABI / kernels
@pytest.mark.parametrize('n,d,q', [(3, 1, 1), (127, 2, 2), (128, 2, 2), (257, 64, 1), (1000, 3, 9)])
def test_edge_shapes(n, d, q):
    """Tiny n, n at / just over a block boundary, the maximum input dimension, q not a multiple of 8."""
    rng = np.random.default_rng(n + d)
    p = max(q, 3)
    x = rng.uniform(0, 1, (n, d))
    y = np.sin(x @ rng.standard_normal((d, p))).T + 0.1 * rng.standard_normal((p, n))
    m, o = _pair(x, y, q=q, submethod='full')
    _check_loss_grad(m, o, o.neglpost_chol)
    x0 = rng.uniform(0, 1, (5, d))
    for a, b in zip(m.predict(x0), o.predict(torch.as_tensor(x0))):
        assert rel(a, b) < PRED_TOL


def test_input_dimension_limit_is_reported():
    rng = np.random.default_rng(0)
    x = rng.uniform(0, 1, (40, 65)); y = rng.standard_normal((3, 40))
    m = LCGP(y=y, x=x, q=2)
    with pytest.raises(ValueError, match='d <= 64'):
        m.loss()
    L = _cabi.lib()
    assert L.lcgp_kernel_matrix(1, 4, 1, 4, 65, 1, 1, 1, 0, 1, None) == -2      # LCGP_E_DIM before any dereference


def test_nan_input_is_reported_not_silently_used():
    x, y = make_full_data(seed=3, n=60, p=3, d=2)
    m = LCGP(y=y, x=x, q=2)
    m.lLmb0.unconstrained.data[0] = float('nan')
    with pytest.raises(RuntimeError, match='Cholesky failed'):
        m.loss()


# ---------------------------------------------------------------- parity at BASELINE's full matrix size
def test_parity_at_config4_matrix_size_q2_p8():
    """n = 8000 unique inputs, d = 10 (config 4's matrix size, 63 blocks; q = 2 latents and p = 8 outputs instead of
    the named q = 32, p = 2000, so that the
    oracle's autograd through two 8000 x 8000 Choleskys stays within a minute and ~30 GB of host memory)."""
    import psutil
    if psutil.virtual_memory().available < 60e9:
        pytest.skip('needs ~30 GB of host memory for the oracle autograd graph')
    torch.set_num_threads(os.cpu_count() or 1)
    x, y, x0, _ = synthetic.latent_mixture(n=8000, d=10, p=8, q_true=4, seed=8000, rep_choices=(1, 2, 3), n0=16)
    m = LCGP(y=y, x=x, q=2, submethod='rep')
    o = O.LCGPOracle(y=y, x=x, q=2, submethod='rep', skip_xnorm=True)
    assert int(m.n) == 8000
    f, g = m.loss_and_grad()
    fo, go = o.loss_and_grad()
    assert abs(f - fo) <= NLL_TOL * abs(fo), (f, fo)
    assert np.max(np.abs(g - go)) <= GRAD_TOL * np.max(np.abs(go))
    yp, ypv, ycv = m.predict(x0)
    ypo, ypvo, ycvo = o.predict(torch.as_tensor(x0))
    assert rel(yp, ypo) < PRED_TOL and rel(ypv, ypvo) < PRED_TOL and rel(ycv, ycvo) < PRED_TOL
    # a point like the end of a fit: large kernel variance, nugget at its lower bound, short length-scales
    lL = m.lLmb.numpy() * 0.35
    m.lLmb.assign(lL); m.lLmb0.assign([25.0, 40.0]); m.lnugGPs.assign([1.2e-7, 1.2e-7]); m.lsigma2s.assign(m.lsigma2s.numpy() - 3.0)
    o.set_constrained(lL, [25.0, 40.0], m.lsigma2s.numpy(), [1.2e-7, 1.2e-7])
    f, g = m.loss_and_grad()
    fo, go = o.loss_and_grad()
    assert abs(f - fo) <= NLL_TOL * abs(fo), (f, fo)
    assert np.max(np.abs(g - go)) <= GRAD_TOL * np.max(np.abs(go))


def test_workspace_guard_bands_untouched():
    """compute-sanitizer is closed on this pool, so out-of-bounds writes are hunted with guard bands:
    the caller-owned workspace and the prediction scratch sit inside larger buffers filled with a
    sentinel; after objective+gradient and predict the bands must be intact (both staging engines write
    only inside the sizes lcgp_workspace_bytes / lcgp_predict_scratch_bytes report)."""
    x, y, xu = make_ragged_rep_data(seed=12, n_unique=300, p=5, d=3)
    m = LCGP(y=y, x=x, q=3, submethod='rep')
    eng = m.engine
    guard = 4096
    sentinel = 12345.6789
    nws = eng.ws_bytes // 8
    big = torch.full((nws + 2 * guard,), sentinel, dtype=DT, device=eng.device)
    eng.ws = big[guard:guard + nws]
    f, g = m.loss_and_grad()
    need = int(eng.lib.lcgp_predict_scratch_bytes(eng.n, eng.q_loc, 200)) // 8
    bigs = torch.full((need + 2 * guard,), sentinel, dtype=DT, device=eng.device)
    eng._scratch = bigs[guard:guard + need]
    out = m.predict(np.random.default_rng(0).uniform(0, 1, (200, 3)))
    torch.cuda.synchronize()
    for buf, n_in in ((big, nws), (bigs, need)):
        assert bool((buf[:guard] == sentinel).all()) and bool((buf[guard + n_in:] == sentinel).all())
    o = O.LCGPOracle(y=y, x=x, q=3, submethod='rep')
    fo, go = o.loss_and_grad()
    assert abs(f - fo) <= NLL_TOL * abs(fo) and np.max(np.abs(g - go)) <= GRAD_TOL * np.max(np.abs(go))
    assert torch.isfinite(out[0]).all()
