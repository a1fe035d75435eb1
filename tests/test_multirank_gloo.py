"""World-size-2 test of the sharded path's HOST logic on CPU ranks (gloo): latent partition,
flat-vector assembly, the single all-reduce per evaluation, prediction gather and lock-step fitting.
The per-latent arithmetic is supplied by tests/helpers.OracleEngine (the CUDA engine needs a GPU)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    import torch.distributed as dist
    torch.set_num_threads(1)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from lcgp_b200 import LCGP
        from helpers import OracleEngine, make_ragged_rep_data, move_params
        x, y, xu = make_ragged_rep_data(seed=1, n_unique=40, p=5, d=2)
        m = LCGP(y=y, x=x, q=q, submethod='rep', diag_error_structure=[2, 3], engine_factory=OracleEngine)
        move_params(m)
        f, g = m.loss_and_grad()
        x0 = np.random.default_rng(3).uniform(0, 1, (9, 2))
        yp, ypv, ycv = m.predict(x0)
        cinv = m.CinvMs.numpy().copy()
        tk = m.Tks[q - 1].numpy().copy()       # dense operator of the LAST latent: rebuilt by its owner, broadcast to all
        m.fit(maxiter=5)                      # also invalidates the cached aux quantities
        assert bool(torch.isnan(m.CinvMs).all())
        ret[rank] = dict(local=m._local_idx.tolist(), f=f, g=g, yp=yp.numpy(), ypv=ypv.numpy(),
                         fitted=m._flat_get(), cinv=cinv, tk=tk)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('q,world', [(3, 2), (1, 2), (1, 3)])     # (1, 3): a rank index beyond q owns nothing
def test_sharded_objective_gradient_predict_fit(q, world):
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from helpers import make_ragged_rep_data, move_params
    from oracle.lcgp_oracle import LCGPOracle
    from test_oracle_selfcheck import _Shim
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, q, ret), nprocs=world, join=True)
    r0, r1 = ret[0], ret[1]
    for rk in range(world):                     # round-robin ownership; ranks >= q own nothing
        assert ret[rk]['local'] == list(range(q))[rk::world]
        assert ret[rk]['f'] == r0['f'] and np.array_equal(ret[rk]['fitted'], r0['fitted'])
    # every rank holds identical (all-reduced) values -> replicated host optimizer stays in lock-step
    assert r0['f'] == r1['f'] and np.array_equal(r0['g'], r1['g'])
    assert np.array_equal(r0['fitted'], r1['fitted'])
    # ... and they equal the unsharded oracle
    x, y, _ = make_ragged_rep_data(seed=1, n_unique=40, p=5, d=2)
    o = LCGPOracle(y=y, x=x, q=q, submethod='rep', diag_error_structure=[2, 3])
    move_params(_Shim(o))
    fo, go = o.loss_and_grad()
    assert abs(r0['f'] - fo) <= 1e-12 * abs(fo)
    assert np.max(np.abs(r0['g'] - go)) <= 1e-9 * np.max(np.abs(go))
    x0 = np.random.default_rng(3).uniform(0, 1, (9, 2))
    ypo, ypvo, _ = o.predict(torch.as_tensor(x0))
    assert np.max(np.abs(r0['yp'] - ypo.numpy())) <= 1e-8 * np.max(np.abs(ypo.numpy()))
    assert np.max(np.abs(r0['ypv'] - ypvo.numpy())) <= 1e-8 * np.max(np.abs(ypvo.numpy()))
    assert np.max(np.abs(r0['cinv'] - o.CinvMs.numpy())) <= 1e-8 * np.max(np.abs(o.CinvMs.numpy()))
    tko = o.Tks[q - 1].numpy()                  # oracle in the stable form (SURVEY B-4)
    for rk in range(world):
        assert np.max(np.abs(ret[rk]['tk'] - tko)) <= 1e-8 * np.max(np.abs(tko))
