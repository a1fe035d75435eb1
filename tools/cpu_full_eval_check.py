"""Checks the x q extrapolation of bench.py's reference arm ONCE: times the oracle (torch-float64 CPU restatement of
lcgp.py:554-630, forward + autograd backward) latent by latent over ALL q latents of a configuration and compares the
total with q x (time of the first latent).  CPU only; ~30 GB of host memory at config 4.

    python tools/cpu_full_eval_check.py cfg4_rep > profiles/r2_cpu_full_eval_check.txt"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from lcgp_b200 import synthetic
from oracle.lcgp_oracle import LCGPOracle

cfg = sys.argv[1] if len(sys.argv) > 1 else 'cfg4_rep'
threads = os.cpu_count() or 1
torch.set_num_threads(threads)
x, y, _, _, mk = synthetic.make_config(cfg)
o = LCGPOracle(y=y, x=x, skip_xnorm=True, **mk)
q = int(o.q)
ts, total_f = [], 0.0
gsum = None
for k in range(q):
    fn = (lambda: o.neglpost_rep(latents=[k])) if mk['submethod'] == 'rep' else (lambda: o.neglpost_chol(latents=[k]))
    t0 = time.time()
    f, g = o.loss_and_grad(fn)
    ts.append(time.time() - t0)
    gsum = g if gsum is None else gsum + g
    print(f'latent {k:2d}: {ts[-1]:7.2f} s', flush=True)
ts = np.array(ts)
print(f'{cfg}: n={int(o.n)} q={q} on {threads} host threads')
print(f'sum over all {q} latents      : {ts.sum():8.1f} s')
print(f'q x first latent              : {q * ts[0]:8.1f} s   (ratio total / extrapolated = {ts.sum() / (q * ts[0]):.3f})')
print(f'q x median latent             : {q * np.median(ts):8.1f} s   (ratio {ts.sum() / (q * np.median(ts)):.3f})')
print(f'per-latent min / median / max : {ts.min():.2f} / {np.median(ts):.2f} / {ts.max():.2f} s')
