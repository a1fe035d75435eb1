"""Host-side behaviour of the drop-in LCGP object on a CPU-only box: the assertions of the reference
test-suite (src/lcgp/tests/test_initialize.py, test_rep.py, test_standardization.py,
test_coverage_gaps.py, test_verification.py) that do not need an objective evaluation, plus
element-wise agreement of the preprocessing with the oracle."""
import numpy as np
import pytest
import torch

from lcgp_b200 import LCGP, evaluation, synthetic
from lcgp_b200.parameter import Parameter, SoftClip
from oracle.lcgp_oracle import LCGPOracle
from helpers import make_full_data, make_ragged_rep_data, make_rep_data


def test_y_must_be_two_dimensional():          # test_initialize.py:9-13
    x = np.linspace(0, 1, 40)
    with pytest.raises(AssertionError):
        LCGP(y=np.sin(x), x=x)


def test_default_q_is_p_and_repr():            # test_initialize.py:15-29
    x, y = make_full_data(n=30, p=3, d=2)
    m = LCGP(y=y, x=x)
    assert m.q == 3 and m.phi.shape == (3, 3)
    assert 'number of latent components:\t3' in repr(m)


@pytest.mark.parametrize('es,ok', [([2, 1], True), ([1, 1, 1], True), (None, True), ([1, 2], True),
                                   ([1, 1], False), ([0, 1, 1], False), ([2, 2], False)])
def test_diag_error_structure(es, ok):         # test_initialize.py:31-42
    x, y = make_full_data(n=30, p=3, d=2)
    if ok:
        m = LCGP(y=y, x=x, diag_error_structure=es)
        assert m.lsigma2s.numpy().shape == (len(es) if es else 3,)
        assert m.get_param()[2].shape == (3,)
    else:
        with pytest.raises(AssertionError):
            LCGP(y=y, x=x, diag_error_structure=es)


def test_q_and_var_threshold_exclusive():      # test_initialize.py:50-54
    x, y = make_full_data(n=30, p=3, d=2)
    with pytest.raises(ValueError):
        LCGP(y=y, x=x, q=2, var_threshold=0.9)
    m = LCGP(y=y, x=x, var_threshold=0.9)
    assert 1 <= m.q <= 3


def test_shape_mismatch_and_bad_submethod():   # test_initialize.py:61-65, test_training.py:21-26
    x, y = make_full_data(n=30, p=3, d=2)
    with pytest.raises(AssertionError):
        LCGP(y=y[:, :-1], x=x)
    with pytest.raises(ValueError):
        LCGP(y=y, x=x, submethod='bogus')


@pytest.mark.parametrize('robust', [True, False])
def test_standardisation_roundtrip(robust):    # test_standardization.py, test_verification.py:37-87
    x, y = make_full_data(n=40, p=4, d=3)
    m = LCGP(y=y, x=x, robust_mean=robust)
    assert float(m.x.min()) == 0.0 and float(m.x.max()) == 1.0 and m.x.shape == (40, 3)
    assert torch.all(m.xnorm > 0)
    assert m.ymean.shape == (4, 1) and m.ystd.shape == (4, 1)
    assert float((m.tx_y(m.y) - m.y_orig).abs().max()) < 1e-10
    assert float((m.tx_x(m.x) - m.x_orig).abs().max()) < 1e-10
    xs, xmin, xmax, xo, xnorm = LCGP.init_standard_x(torch.as_tensor(x))
    assert xs.shape == (40, 3) and torch.all(xnorm > 0)


def test_replication_structures():             # test_rep.py:33-138
    x, y, xu = make_rep_data(seed=99, n_unique=10, p=4, d=2, reps=3)
    m = LCGP(y=y, x=x, submethod='rep')
    assert int(m.n.numpy()) == 10 and m.ybar.shape == (4, 10) and m._rep_initialized
    _, inv, cnt = np.unique(x, axis=0, return_inverse=True, return_counts=True)
    inv = np.asarray(inv).reshape(-1)
    for i in range(10):
        np.testing.assert_allclose(m.ybar.numpy()[:, i], y[:, inv == i].mean(axis=1), atol=1e-10)
    assert np.all(m.r.numpy() == 3)
    assert float(m.x_unique_s.min()) >= 0 and float(m.x_unique_s.max()) <= 1
    np.testing.assert_allclose(m.R.numpy(), np.diag(m.r.numpy().astype(float)))
    for a in ['x_unique', 'x_unique_s', 'group_ids', 'r', 'ybar', 'ybar_s', 'ybar_mean', 'ybar_std']:
        assert hasattr(m, a)


def test_preprocess_tuple_and_helpers():       # test_coverage_gaps.py:20-126
    x, y, _ = make_rep_data(n_unique=15, p=3, d=2, reps=4)
    m = LCGP(y=y, x=x, submethod='rep', robust_mean=False)
    c, s = m._compute_center_spread_tf(m.ybar)
    np.testing.assert_allclose(c.numpy(), m.ybar.numpy().mean(axis=1, keepdims=True))
    np.testing.assert_allclose(s.numpy(), m.ybar.numpy().std(axis=1, keepdims=True))
    res = m.preprocess(x_raw=x, y_raw=y)
    assert len(res) == 12
    assert int(res[9].numpy()) == 15 and int(res[10].numpy()) == 2 and int(res[11].numpy()) == 3
    assert res[0].shape == (15, 2) and res[5].shape == (3, 15) and res[6].shape == (3, 15) and res[4].shape == (15, 15)
    assert m.preprocess()[9].numpy() == 15
    m._rep_initialized = False
    calls = []
    orig = m.preprocess
    m.preprocess = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
    m._ensure_replication()
    assert calls == [1] and m._rep_initialized
    m._ensure_replication()
    assert calls == [1]
    m2 = LCGP(y=y, x=x, submethod='rep', rep_standardize_ybar=False)
    np.testing.assert_allclose(m2._get_phi_input().numpy(), m2.ybar.numpy())
    del m2.ybar_s, m2.ybar
    np.testing.assert_allclose(m2._get_phi_input().numpy(), m2.y.numpy())


def test_phi_reconstructs_ybar_when_q_equals_p():   # test_verification.py:89-136
    x, y, _ = make_ragged_rep_data(seed=0, n_unique=50, p=3, d=2)
    m = LCGP(y=y, x=x, submethod='rep')
    # phi = U sqrt(n)/s  =>  (phi D^-1) g = U U^T Y = Y  when q = p
    rec = (m.phi / m.diag_D) @ m.g
    assert float((rec - m.ybar_s).abs().max()) < 1e-8


@pytest.mark.parametrize('sub,robust,std', [('rep', True, True), ('rep', False, False), ('full', True, True)])
def test_preprocessing_matches_oracle(sub, robust, std):
    x, y, _ = make_ragged_rep_data(seed=5, n_unique=45, p=5, d=3)
    kw = dict(q=3, submethod=sub, robust_mean=robust, rep_standardize_ybar=std, diag_error_structure=[2, 3])
    m, o = LCGP(y=y, x=x, **kw), LCGPOracle(y=y, x=x, **kw)
    names = ['x', 'x_min', 'x_max', 'g', 'diag_D', 'xnorm']
    names += ['x_unique', 'x_unique_s', 'ybar', 'ybar_s', 'ybar_mean', 'ybar_std'] if sub == 'rep' else ['y', 'ymean', 'ystd']
    for a in names:
        assert float((getattr(m, a) - getattr(o, a)).abs().max()) < 1e-12, a
    assert float((m.phi.abs() - o.phi.abs()).abs().max()) < 1e-12          # SVD sign gauge
    for a, b in zip(m.trainable_variables, o.trainable_variables):
        assert float((a - b).abs().max()) < 1e-12
    for a, b in zip(m.get_param(), o.get_param()):
        assert float((a - b.detach()).abs().max()) < 1e-12


def test_replicate_means_follow_the_reference_loop():
    """_compute_ybar_np (segmented sum over group-sorted columns) == the reference's per-group loop
    `ybar[:, i] = y[:, inverse == i].mean(axis=1)` (lcgp.py:358-367) on ragged replication, to summation-order
    rounding (numpy's mean adds groups of >= 8 replicates with 8 interleaved partial sums, the segmented sum
    left to right); identical bits for groups of fewer than 8."""
    x, y, xu = make_ragged_rep_data(seed=4, n_unique=57, p=6, d=2)
    xr, inv, cnt = LCGP._group_unique_rows_np(np.asarray(x))
    n = xr.shape[0]
    yr = np.asarray(y)
    loop = np.stack([yr[:, inv == i].mean(axis=1) for i in range(n)], axis=1)
    got = LCGP._compute_ybar_np(yr, inv, n)
    assert got.shape == (6, n) and np.allclose(got, loop, rtol=4e-16 * cnt.max(), atol=1e-300)
    heavy = make_ragged_rep_data(seed=5, n_unique=9, p=3, d=1)
    xh, ih, ch = LCGP._group_unique_rows_np(np.asarray(np.repeat(heavy[0], 5, axis=0)))      # groups of >= 8 too
    yh = np.random.default_rng(0).normal(size=(3, ih.size))
    lh = np.stack([yh[:, ih == i].mean(axis=1) for i in range(xh.shape[0])], axis=1)
    assert ch.max() >= 8 and np.allclose(LCGP._compute_ybar_np(yh, ih, xh.shape[0]), lh, rtol=4e-16 * ch.max(), atol=1e-300)
    small = cnt < 8
    assert small.any() and np.array_equal(got[:, small], loop[:, small])
    order, off = LCGP._segments(inv, n)
    assert off[0] == 0 and off[-1] == yr.shape[1] and np.array_equal(np.diff(off), cnt)
    for i in (0, n // 2, n - 1):          # stable: original column order inside each group
        assert np.array_equal(order[off[i]:off[i + 1]], np.nonzero(inv == i)[0])


def test_xnorm_sort_formula_matches_dense_definition():
    rng = np.random.default_rng(0)
    x = torch.as_tensor(np.round(rng.uniform(0, 1, (70, 2)), 1))    # many ties
    dense = torch.zeros(2, dtype=torch.float64)
    for j in range(2):
        dmat = (x[:, j][:, None] - x[:, j][None, :]).abs()
        dense[j] = dmat[dmat > 0].mean()
    assert float((LCGP._mean_positive_distance(x) - dense).abs().max()) < 1e-13


def test_parameter_container():
    p = Parameter(np.array([0.5, 2.0]), SoftClip(1e-6, 1e4), name='t')
    np.testing.assert_allclose(p.numpy(), [0.5, 2.0], rtol=1e-10)
    p.assign([0.25, 3.0])
    np.testing.assert_allclose(np.asarray(p), [0.25, 3.0], rtol=1e-10)
    v = p.value().sum(); v.backward()
    assert p.unconstrained.grad is not None and p.shape == (2,)


def test_no_cuda_means_loud_failure():
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    x, y = make_full_data(n=30, p=3, d=2)
    m = LCGP(y=y, x=x, q=2)
    for call in (m.loss, m.fit, lambda: m.predict(x[:3]), m.neglpost):
        with pytest.raises(RuntimeError, match='no CPU fallback'):
            call()
    from lcgp_b200 import Matern32
    with pytest.raises(AssertionError):                       # test_cov.py:18-23
        Matern32(np.zeros(3), np.zeros(3), 1.0, 1.0, 1e-3)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        Matern32(x, x, np.ones(2), 1.0, 1e-3)
    d = Matern32(x, x, np.ones(2), 2.0, 1e-3, diag_only=True)   # host-only branch, covmat.py:23-29
    np.testing.assert_allclose(d.numpy(), 2.0 * np.ones(30))
    # the entry points added later fail just as loudly
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        m.grad_phi()
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        LCGP(y=y, x=x, q=2, device_preprocess=True)
    from lcgp_b200.model import _fullcov_cuda
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        _fullcov_cuda(torch.ones(2, 3, dtype=torch.float64), torch.ones(2, 4, dtype=torch.float64),
                      torch.ones(3, dtype=torch.float64), torch.ones(3, dtype=torch.float64))
    assert m._dev_prep is False          # constructed on the host pipeline when no CUDA device exists


def test_evaluation_metrics():                 # test_diagnostics.py
    rng = np.random.default_rng(0)
    y = rng.standard_normal((3, 20))
    assert evaluation.rmse(y, y) == 0 and evaluation.normalized_rmse(y, y) == 0
    cov, width = evaluation.intervalstats(y, y + 0.1, np.ones_like(y))
    assert 0 <= cov <= 1 and width > 0
    assert np.isfinite(evaluation.dss(y, y + 0.1, np.ones_like(y), use_diag=True))
    full = np.stack([np.eye(3)] * 20, axis=2)
    assert abs(evaluation.dss(y, y + 0.1, full, use_diag=False) - evaluation.dss(y, y + 0.1, np.ones_like(y), True)) < 1e-12


def test_dss_full_covariance_matches_its_definition():
    """evaluation.dss(use_diag=False) (src/lcgp/evaluation.py:21-49: log det Sigma + r^T Sigma^-1 r through an
    eigen-decomposition, averaged over test points) against a direct solve, for random SPD covariances."""
    rng = np.random.default_rng(3)
    p, n = 4, 9
    y, mu = rng.standard_normal((p, n)), rng.standard_normal((p, n))
    A = rng.standard_normal((n, p, p))
    S = np.einsum('nij,nkj->nik', A, A) + 0.5 * np.eye(p)
    want = np.mean([np.linalg.slogdet(S[i])[1] + (y[:, i] - mu[:, i]) @ np.linalg.solve(S[i], y[:, i] - mu[:, i]) for i in range(n)])
    assert abs(evaluation.dss(y, mu, np.moveaxis(S, 0, 2), use_diag=False) - want) <= 1e-12 * abs(want)
    d = rng.uniform(0.5, 2.0, (p, n))                  # diagonal covariances: both forms agree
    Sd = np.stack([np.diag(d[:, i]) for i in range(n)], axis=2)
    assert abs(evaluation.dss(y, mu, Sd, use_diag=False) - evaluation.dss(y, mu, d, use_diag=True)) < 1e-12


def test_example_harness_interface():
    """LCGPRun / SuperRun (docs/call_model.py:5-86): constructor keywords, attributes, swallowed extras; the model is
    built with the harness' settings.  (train / predict need the CUDA path: tests/test_gpu_parity.py.)"""
    from lcgp_b200 import LCGPRun
    x, y = make_full_data(n=30, p=3, d=2)
    data = dict(xtrain=x, ytrain=y, xtest=x[:5], ytest=y[:, :5], ytrue=y[:, :5])
    run = LCGPRun(runno='r0', data=data, submethod='rep', robust=False, num_latent=2, err_struct=[1, 2],
                  diag_error_structure=[3], robust_mean=True)                 # the last two are swallowed (SURVEY B-10)
    assert run.modelname == 'LCGP' and run.n == 30 and run.num_output == 3 and run.model is None and hasattr(run, 'ytrue')
    assert LCGPRun(runno='r1', data=data).modelname == 'LCGP_robust'
    run.define_model()
    m = run.model
    assert m.submethod == 'rep' and int(m.q) == 2 and m.robust_mean is False and m.diag_error_structure == [1, 2]


def test_synthetic_configs_shapes():
    x, y, x0, y0, mk = synthetic.make_config('cfg5_one')
    assert x.shape == (1024, 6) and y.shape == (64, 1024) and mk['q'] == 8
    x, y, x0, y0, mk = synthetic.make_config('cfg3_rep', n=100)
    assert np.unique(x, axis=0).shape[0] == 100 and y.shape[0] == 500


def test_state_dict_roundtrip(tmp_path):
    x, y = make_full_data(n=30, p=3, d=2)
    a = LCGP(y=y, x=x, q=2)
    a.lLmb.assign(a.lLmb.numpy() * 1.7); a.lsigma2s.assign([-1.0, -2.0, -3.0])
    torch.save(a.state_dict(), tmp_path / 'm.pt')
    b = LCGP(y=y, x=x, q=2).load_state_dict(torch.load(tmp_path / 'm.pt'))
    for pa, pb in zip(a.get_param(), b.get_param()):
        assert torch.equal(pa, pb)
    with pytest.raises(ValueError):
        LCGP(y=y, x=x, q=3).load_state_dict(a.state_dict())
    with pytest.raises(ValueError):
        LCGP(y=y * 2.0 + 1.0, x=x[::-1].copy(), q=2).load_state_dict(a.state_dict())


def test_lbfgsb_machine_equals_scipy_minimize():
    """lcgp_b200.lbfgsb.LbfgsbMachine (the reverse-communication form the lock-step batched fits advance) visits the
    points of scipy.optimize.minimize(method='L-BFGS-B') -- the reference's optimizer, lcgp.py:537-540 -- bit for bit,
    with default options and with an iteration limit."""
    import scipy.optimize as so
    from lcgp_b200.lbfgsb import LbfgsbMachine

    def fg(x):
        return so.rosen(x) + 0.1 * np.sum(np.sin(3 * x)), so.rosen_der(x) + 0.3 * np.cos(3 * x)
    x0 = np.linspace(-1.2, 1.5, 12)
    for opts in ({}, {'maxiter': 7}, {'maxcor': 4, 'gtol': 1e-9, 'ftol': 1e-14}):
        path = []

        def fun(x):
            path.append(x.copy())
            return fg(x)
        r = so.minimize(fun, x0, jac=True, method='L-BFGS-B', options=opts or None)
        m = LbfgsbMachine(x0, **opts)
        path2 = []
        while m.advance():
            path2.append(m.x.copy())
            m.supply(*fg(m.x))
        assert len(path) == len(path2) and all(np.array_equal(a, b) for a, b in zip(path, path2))
        assert (r.nfev, r.nit, bool(r.success)) == (m.nfev, m.nit, m.success)
        assert np.array_equal(r.x, m.x) and r.fun == float(m.f)
    assert not m.advance()                              # stays stopped


def test_toggling_rep_standardize_ybar_rebuilds_the_engine():
    """The reference marks rep_standardize_ybar "can toggle" (lcgp.py:49).  The engine's constant data depend on it, so a
    toggle must drop the engine and the cached predictive quantities: the objective afterwards is the one of a model
    evaluated with that flag, never a mix of the two standardisations."""
    from helpers import OracleEngine, make_ragged_rep_data
    x, y, _ = make_ragged_rep_data(seed=4, n_unique=30, p=3, d=2)
    m = LCGP(y=y, x=x, q=2, submethod='rep', engine_factory=OracleEngine)
    f_std = float(m.loss())
    eng = m.engine
    m.rep_standardize_ybar = False
    assert m._engine is None and bool(torch.isnan(m.CinvMs).all())
    f_raw = float(m.loss())
    assert m.engine is not eng and abs(f_raw - f_std) > 1e-6 * abs(f_std)
    m.rep_standardize_ybar = True
    assert abs(float(m.loss()) - f_std) <= 1e-13 * abs(f_std)
    m.rep_standardize_ybar = True                       # no change: the engine stays
    e2 = m.engine
    m.rep_standardize_ybar = 1
    assert m.engine is e2
