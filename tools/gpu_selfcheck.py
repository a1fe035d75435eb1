"""Stage-by-stage numerical check of the CUDA path against torch CPU float64 (developer tool).

Prints one line per stage with the max error instead of asserting, so a single GPU run shows
where a discrepancy starts.  Usage (on a GPU box):  python tools/gpu_selfcheck.py [n] [d] [q]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from lcgp_b200 import _cabi
from oracle.lcgp_oracle import Matern32 as oracle_matern

DT = torch.float64


def rel(a, b):
    a, b = a.detach().cpu(), b.detach().cpu()
    den = float(b.abs().max())
    return float((a - b).abs().max()) / (den if den > 0 else 1.0)


def stage_check(n=300, d=3, q=2, seed=0):
    L = _cabi.lib()
    dev = torch.device('cuda')
    rng = np.random.default_rng(seed)
    X = torch.as_tensor(rng.uniform(0, 1, (n, d)))
    r = torch.as_tensor(rng.integers(1, 4, n).astype(np.float64))
    sr = torch.sqrt(r)
    ell = torch.as_tensor(rng.uniform(0.3, 1.5, (q, d)))
    s0 = torch.as_tensor(rng.uniform(0.5, 3.0, q))
    lnug = torch.as_tensor(np.exp(rng.uniform(-12, -5, q)))
    D = torch.as_tensor(rng.uniform(0.3, 2.0, q))
    npad = _cabi.padded(n)
    nb = npad // _cabi.NB
    st = _cabi.stream_ptr()
    g = lambda t: t.to(dev).contiguous()
    Xd, srd, elld, s0d, nugd, Dd = g(X), g(sr), g(ell), g(s0), g(lnug), g(D)

    # --- kernel matrix
    C_ref = oracle_matern(X, X, ell[0], s0[0], lnug[0])
    out = torch.empty((n, n), dtype=DT, device=dev)
    rc = L.lcgp_kernel_matrix(Xd.data_ptr(), n, Xd.data_ptr(), n, d, elld[0].contiguous().data_ptr(), s0d[0:1].data_ptr(),
                              nugd[0:1].data_ptr(), 1, out.data_ptr(), st)
    torch.cuda.synchronize()
    print(f'[kernel_matrix] rc={rc} rel err {rel(out, C_ref):.3e}')

    # --- build A
    F = torch.full((q, npad, npad), float('nan'), dtype=DT, device=dev)
    rc = L.lcgp_build_A(Xd.data_ptr(), srd.data_ptr(), n, d, elld.data_ptr(), s0d.data_ptr(), nugd.data_ptr(),
                        Dd.data_ptr(), q, F.data_ptr(), npad, st)
    torch.cuda.synchronize()
    A_ref = []
    for k in range(q):
        Ck = oracle_matern(X, X, ell[k], s0[k], lnug[k])
        A = torch.eye(npad, dtype=DT)
        A[:n, :n] = torch.eye(n, dtype=DT) + D[k] * ((Ck * sr[None, :]) * sr[:, None])
        A_ref.append(A)
    A_ref = torch.stack(A_ref)
    low = torch.tril(torch.ones(npad, npad, dtype=torch.bool))
    Fc = F.cpu()
    print(f'[build_A] rc={rc} lower-tri rel err {rel(Fc[:, low], A_ref[:, low]):.3e}  nan in lower: {bool(torch.isnan(Fc[:, low]).any())}')

    # --- potrf
    DL = torch.zeros((q, nb, 128, 128), dtype=DT, device=dev)
    DU = torch.zeros((q, nb, 128, 128), dtype=DT, device=dev)
    ldp = torch.zeros((q, nb), dtype=DT, device=dev)
    info = torch.zeros(q, dtype=torch.int32, device=dev)
    sb = int(L.lcgp_potrf_scratch_bytes(npad, q)); pscr = torch.zeros(sb // 4, dtype=torch.int32, device=F.device)
    rc = L.lcgp_potrf_batched(F.data_ptr(), npad, q, DL.data_ptr(), DU.data_ptr(), ldp.data_ptr(), info.data_ptr(), pscr.data_ptr(), sb, st)
    torch.cuda.synchronize()
    L_ref = torch.linalg.cholesky(A_ref)
    Fc = F.cpu()
    print(f'[potrf] rc={rc} info={info.cpu().tolist()} L rel err {rel(Fc[:, low], L_ref[:, low]):.3e} '
          f'logdet rel err {rel(ldp.sum(1), torch.log(torch.diagonal(L_ref, dim1=1, dim2=2)).sum(1)):.3e}')
    Linv_ref = torch.linalg.solve_triangular(L_ref, torch.eye(npad, dtype=DT).expand_as(L_ref), upper=False)
    for b in range(nb):
        sl = slice(b * 128, (b + 1) * 128)
        e1 = rel(DL[:, b], Linv_ref[:, sl, sl])
        e2 = rel(DU[:, b], Linv_ref[:, sl, sl].transpose(1, 2))
        if b < 3 or b == nb - 1:
            print(f'   diag block {b}: DL rel err {e1:.3e}  DU rel err {e2:.3e}')

    # --- trtri
    sb = int(L.lcgp_trtri_scratch_bytes(npad, q))
    scratch = torch.empty(max(sb // 8, 1), dtype=DT, device=dev)
    rc = L.lcgp_trtri_batched(F.data_ptr(), npad, q, DL.data_ptr(), DU.data_ptr(), scratch.data_ptr(), sb, st)
    torch.cuda.synchronize()
    Fc = F.cpu()
    U_ref = Linv_ref.transpose(1, 2)
    blk_upper = torch.zeros(npad, npad, dtype=torch.bool)
    for I in range(nb):
        for J in range(I + 1, nb):
            blk_upper[I * 128:(I + 1) * 128, J * 128:(J + 1) * 128] = True
    if blk_upper.any():
        print(f'[trtri] rc={rc} strictly-upper blocks rel err {rel(Fc[:, blk_upper], U_ref[:, blk_upper]):.3e}')
    else:
        print(f'[trtri] rc={rc} (single block, nothing to do)')


def model_check(n_unique=120, d=3, p=5, q=3, submethod='rep', seed=1):
    from lcgp_b200 import LCGP
    from oracle.lcgp_oracle import LCGPOracle
    rng = np.random.default_rng(seed)
    xu = rng.uniform(0, 1, (n_unique, d))
    if submethod == 'rep':
        r = rng.integers(1, 4, n_unique)
        x = np.repeat(xu, r, axis=0)
    else:
        x = xu
    y = np.sin(x @ rng.standard_normal((d, p))).T + 0.1 * rng.standard_normal((p, x.shape[0]))
    m = LCGP(y=y, x=x, q=q, submethod=submethod)
    o = LCGPOracle(y=y, x=x, q=q, submethod=submethod)
    # move off the init point
    lL = m.lLmb.numpy() * rng.uniform(0.5, 2.0, (q, d))
    l0 = rng.uniform(0.5, 3.0, q)
    ls = m.lsigma2s.numpy() + rng.normal(0, 0.3, p)
    ln = np.exp(rng.uniform(-12, -5, q))
    m.lLmb.assign(lL); m.lLmb0.assign(l0); m.lsigma2s.assign(ls); m.lnugGPs.assign(ln)
    o.set_constrained(lL, l0, ls, ln)
    t = time.time()
    f, g = m.loss_and_grad()
    t1 = time.time() - t
    fo, go = o.loss_and_grad(o.neglpost_chol if submethod == 'full' else None)
    print(f'[model {submethod} n={n_unique} d={d} p={p} q={q}] nll {f:.15g} oracle {fo:.15g} rel {abs(f - fo) / abs(fo):.3e}  '
          f'grad rel {np.max(np.abs(g - go)) / np.max(np.abs(go)):.3e}  ({t1 * 1e3:.1f} ms first eval)')
    o1 = 0
    for name, v in zip(['lLmb', 'lLmb0', 'lnug', 'lsig'], m.trainable_variables):
        k = v.numel()
        den = np.max(np.abs(go[o1:o1 + k]))
        print(f'    grad block {name}: rel {np.max(np.abs(g[o1:o1 + k] - go[o1:o1 + k])) / (den if den > 0 else 1):.3e}')
        o1 += k
    x0 = rng.uniform(0, 1, (37, d))
    res = m.predict(x0)
    reso = o.predict(x0)
    for nm, a, b in zip(['ypred', 'ypredvar', 'yconfvar'], res, reso):
        print(f'    predict {nm}: rel {rel(a, b):.3e}')
    print(f'    CinvMs rel {rel(m.CinvMs, o.CinvMs):.3e}')
    xt = m.x_unique if submethod == 'rep' else m.x_orig
    res = m.predict(xt)
    reso = o.predict(xt)
    print(f'    predict at training inputs (nugget quirk): ypred rel {rel(res[0], reso[0]):.3e} var rel {rel(res[1], reso[1]):.3e}')


if __name__ == '__main__':
    args = [int(a) for a in sys.argv[1:]]
    print(_cabi.lib().lcgp_version().decode(), '|', torch.cuda.get_device_name(0))
    stage_check(40, 1, 2)
    stage_check(300, 3, 2)
    stage_check(*(args or [1000, 8, 2]))
    model_check(40, 1, 3, 3, 'rep')
    model_check(120, 3, 5, 3, 'rep')
    model_check(150, 2, 4, 4, 'full')
    model_check(700, 5, 12, 4, 'rep')
