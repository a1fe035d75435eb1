"""Worker of tests/test_nccl_sharded.py (one process per GPU, NCCL): the sharded evaluation path of LCGP
(lcgp.py:605-624's q-loop split over ranks + one all-reduce) against the unsharded CUDA path and the CPU
oracle on the same data.  Launched with torch.distributed.run; exits non-zero on any mismatch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import numpy as np
import torch
import torch.distributed as dist

from lcgp_b200 import LCGP, synthetic
from oracle.lcgp_oracle import LCGPOracle
from helpers import move_params


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def main():
    rank, world, lr = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(lr)
    dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
    torch.set_num_threads(2)
    ok = True
    cases = {
        'rep3d q=3': (*synthetic.rep3d(), dict(q=3, submethod='rep')),
        'rep n=1000 q=5': (*synthetic.latent_mixture(n=1000, d=6, p=40, q_true=5, seed=1, rep_choices=(1, 2), n0=50)[:3],
                           dict(q=5, submethod='rep')),
        'full n=400 q=1 (a rank owns no latent)': (*synthetic.latent_mixture(n=400, d=3, p=6, q_true=2, seed=2, rep_choices=None,
                                                                             n0=30)[:3], dict(q=1, submethod='full')),
    }
    for name, (x, y, x0, mk) in cases.items():
        ms = LCGP(y=y, x=x, shard=True, **mk)          # latents sharded over the ranks
        mu = LCGP(y=y, x=x, shard=False, **mk)         # every rank also computes everything locally
        o = LCGPOracle(y=y, x=x, **mk)
        move_params(ms, o); move_params(mu)
        fn = o.neglpost_chol if mk['submethod'] == 'full' else None
        fs, gs = ms.loss_and_grad(); fu, gu = mu.loss_and_grad(); fo, go = o.loss_and_grad(fn)
        e_su = max(abs(fs - fu) / abs(fu), rel(gs, gu))
        e_f, e_g = abs(fs - fo) / abs(fo), rel(gs, go)
        ps, po = ms.predict(x0), o.predict(torch.as_tensor(x0))
        e_p = max(rel(a.numpy(), b.numpy()) for a, b in zip(ps, po))
        ms.fit(maxiter=8); mu.fit(maxiter=8)
        e_fit = rel(ms._flat_get(), mu._flat_get())
        flat = torch.tensor(ms._flat_get(), device='cuda'); ref = flat.clone(); dist.broadcast(ref, 0)
        lock = bool(torch.equal(flat, ref))             # bit-identical parameters on every rank after the fit
        good = e_su < 1e-12 and e_f < 1e-10 and e_g < 1e-8 and e_p < 1e-6 and e_fit < 1e-6 and lock
        ok &= good
        if rank == 0:
            print(f'[{name}] world={world} local={ms._local_idx.tolist()} sharded-vs-unsharded {e_su:.1e} vs oracle: '
                  f'loss {e_f:.1e} grad {e_g:.1e} predict {e_p:.1e}; fit {e_fit:.1e} lock-step {lock} -> '
                  f'{"OK" if good else "FAIL"}', flush=True)
    # a Cholesky failure on ONE rank's latent must raise on EVERY rank (no hang in the next all-reduce)
    x, y, x0 = synthetic.rep3d()
    m = LCGP(y=y, x=x, q=3, submethod='rep', shard=True)
    m.lLmb0.unconstrained.data[1] = float('nan')        # latent 1 lives on rank 1 % world
    try:
        m.loss()
        raised = False
    except RuntimeError as ex:
        raised = 'Cholesky failed' in str(ex)
    ok &= raised
    m.lLmb0.assign(np.ones(3))
    ok &= bool(np.isfinite(float(m.loss())))            # and the job keeps working afterwards
    if rank == 0:
        print(f'[collective failure] raised on this rank: {raised}', flush=True)
    flag = torch.tensor([int(ok)], device='cuda')
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print('NCCL_SHARDED', 'PASS' if int(flag) else 'FAIL', flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == '__main__':
    main()
