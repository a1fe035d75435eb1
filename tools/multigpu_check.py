"""Sharded (one rank per GPU, NCCL) vs unsharded results on the same data: objective, gradient, fit, predict.
Run:  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/multigpu_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from lcgp_b200 import LCGP, synthetic

rank, world, lr = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
torch.set_num_threads(2)
ok = True
for name, (x, y, x0, mk) in {
    'rep3d q=3': (*synthetic.rep3d(), dict(q=3, submethod='rep')),
    'cfg3-like n=1000 q=5': (*synthetic.latent_mixture(n=1000, d=6, p=40, q_true=5, seed=1, rep_choices=(1, 2), n0=50)[:3], dict(q=5, submethod='rep')),
    'full n=400 q=2 (q < world possible)': (*synthetic.latent_mixture(n=400, d=3, p=6, q_true=2, seed=2, rep_choices=None, n0=30)[:3], dict(q=2, submethod='full')),
}.items():
    ms = LCGP(y=y, x=x, shard=True, **mk)          # sharded over the ranks
    mu = LCGP(y=y, x=x, shard=False, **mk)         # every rank also computes the whole thing locally
    fs, gs = ms.loss_and_grad(); fu, gu = mu.loss_and_grad()
    e_f = abs(fs - fu) / abs(fu); e_g = np.max(np.abs(gs - gu)) / np.max(np.abs(gu))
    ps, pu = ms.predict(x0), mu.predict(x0)
    e_p = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(ps, pu))
    ms.fit(maxiter=8); mu.fit(maxiter=8)
    e_fit = float(np.max(np.abs(ms._flat_get() - mu._flat_get())) / np.max(np.abs(mu._flat_get())))
    flat = torch.tensor(ms._flat_get(), device='cuda'); ref = flat.clone(); dist.broadcast(ref, 0)
    lock = bool(torch.equal(flat, ref))             # all ranks hold bit-identical parameters after the fit
    good = e_f < 1e-12 and e_g < 1e-10 and e_p < 1e-9 and e_fit < 1e-6 and lock
    ok &= good
    if rank == 0:
        print(f'[{name}] world={world} local latents={ms._local_idx.tolist()} loss rel {e_f:.2e} grad rel {e_g:.2e} '
              f'predict rel {e_p:.2e} fitted-params rel {e_fit:.2e} lock-step {lock} -> {"OK" if good else "FAIL"}')
flag = torch.tensor([int(ok)], device='cuda'); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print('MULTIGPU_CHECK', 'PASS' if int(flag) else 'FAIL')
dist.destroy_process_group()
sys.exit(0 if int(flag) else 1)
