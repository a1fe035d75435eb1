"""Times lcgp_potrf_batched in isolation (developer tool): np=128 isolates the diagonal-block kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lcgp_b200 import _cabi
L = _cabi.lib(); dev = torch.device('cuda'); DT = torch.float64

def run(npad, batch, reps=20, check=False):
    nb = npad // 128
    A = torch.randn(batch, npad, npad, dtype=DT, device=dev)
    A = A @ A.transpose(1, 2) / npad + 2 * torch.eye(npad, dtype=DT, device=dev)
    DLb = torch.zeros((batch, nb, 128, 128), dtype=DT, device=dev); DUb = torch.zeros_like(DLb)
    info = torch.zeros(batch, dtype=torch.int32, device=dev)
    st = _cabi.stream_ptr()
    sb = int(L.lcgp_potrf_scratch_bytes(npad, batch))
    scr = torch.zeros(max(sb // 4, 1), dtype=torch.int32, device=dev)
    legacy = os.environ.get('LCGP_POTRF') == 'panels'
    ts = []
    for r in range(reps):
        F = A.clone()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.lcgp_potrf_batched(F.data_ptr(), npad, batch, DLb.data_ptr(), DUb.data_ptr(), None, info.data_ptr(),
                             None if legacy else scr.data_ptr(), 0 if legacy else sb, st)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    if check and not os.environ.get('LCGP_DIAG_MODE'):
        Lref = torch.linalg.cholesky(A[0])
        err = (torch.tril(F[0]) - Lref).abs().max().item() / Lref.abs().max().item()
        print(f'   check vs torch.linalg.cholesky: max rel err {err:.2e}  info {info.tolist()}')
        L00 = torch.tril(F[0][:128, :128])
        inv_err = (DLb[0, 0] @ L00 - torch.eye(128, dtype=DT, device=dev)).abs().max().item()
        du_err = (DUb[0, 0] - DLb[0, 0].T).abs().max().item()
        print(f'   diagonal-block inverse: |DL L - I| {inv_err:.2e}  |DU - DL^T| {du_err:.2e}')
        assert err < 1e-12 and inv_err < 1e-11 and du_err == 0.0
    ts.sort()
    flops = batch * npad ** 3 / 3
    print(f'np={npad:5d} batch={batch:3d}  median {ts[len(ts)//2]:9.1f} us  min {ts[0]:9.1f} us   {flops / (ts[len(ts)//2] * 1e-6) / 1e12:6.2f} TF/s')

import os
CASES = [(128, 1)] if os.environ.get("DIAG_ONLY") else [tuple(int(v) for v in c.split('x')) for c in os.environ["CASES"].split(',')] if os.environ.get("CASES") else [(128, 1), (128, 8), (128, 32), (256, 1), (512, 1), (1024, 1), (1024, 8), (2048, 1), (2048, 10), (4096, 4), (8064, 1), (8064, 4)]
for npad, batch in CASES:
    run(npad, batch, check=(npad in (128, 1024, 8064) and batch == 1))
