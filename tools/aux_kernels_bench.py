"""Times the auxiliary kernels at config-4 size with CUDA events (developer tool; JSON line on stdout):
full predictive covariance (p=2000, q=32, n0=64), replicate means / median / MAD / standardise (p=2000, N=16000)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lcgp_b200 import _cabi
L = _cabi.lib(); dev = torch.device('cuda'); DT = torch.float64
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))

def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))

out = {'hbm_peak_gbs': peaks['hbm_gbs']}
g = torch.Generator(device='cuda').manual_seed(0)
for q, p, n0 in [(32, 2000, 64), (3, 2000, 64), (32, 512, 512)]:
    psi = torch.randn(q, p, dtype=DT, device=dev, generator=g); gv = torch.rand(q, n0, dtype=DT, device=dev, generator=g)
    s2 = torch.rand(p, dtype=DT, device=dev, generator=g); sv = torch.rand(p, dtype=DT, device=dev, generator=g) + 0.5
    buf = torch.empty((n0, p, p), dtype=DT, device=dev)
    ms = timed(lambda: L.lcgp_predict_fullcov(psi.data_ptr(), gv.data_ptr(), s2.data_ptr(), sv.data_ptr(), q, p, n0, buf.data_ptr(), _cabi.stream_ptr()))
    gbs = 8.0 * n0 * p * p / (ms * 1e-3) / 1e9
    out[f'fullcov_q{q}_p{p}_n0{n0}'] = {'ms': ms, 'write_gbs': gbs, 'frac_of_hbm': gbs / peaks['hbm_gbs'], 'tflops': 2.0 * q * n0 * p * p / (ms * 1e-3) / 1e12}
p, N, n = 2000, 16000, 8000
y = torch.randn(p, N, dtype=DT, device=dev, generator=g)
inv = np.sort(np.random.default_rng(0).integers(0, n, N)); inv[:n] = np.arange(n); inv = np.random.default_rng(1).permutation(inv)
order = np.argsort(inv, kind='stable').astype(np.int32); off = np.concatenate([[0], np.cumsum(np.bincount(inv, minlength=n))]).astype(np.int32)
od, fd = torch.as_tensor(order).to(dev), torch.as_tensor(off).to(dev)
ybar = torch.empty((p, n), dtype=DT, device=dev)
ms = timed(lambda: L.lcgp_prep_segment_mean(y.data_ptr(), od.data_ptr(), fd.data_ptr(), p, N, n, ybar.data_ptr(), _cabi.stream_ptr()))
out['segment_mean'] = {'ms': ms, 'gbs': 8.0 * p * (N + n) / (ms * 1e-3) / 1e9}
c = torch.empty(p, dtype=DT, device=dev); s = torch.empty(p, dtype=DT, device=dev)
k = int(np.round((n - 1) * 0.5))
ms = timed(lambda: L.lcgp_prep_row_select(ybar.data_ptr(), None, p, n, k, c.data_ptr(), _cabi.stream_ptr()))
out['row_select_median'] = {'ms': ms, 'gbs_one_pass': 8.0 * p * n / (ms * 1e-3) / 1e9, 'note': '8 passes over an L2-resident row per CTA'}
ms = timed(lambda: L.lcgp_prep_row_select(ybar.data_ptr(), c.data_ptr(), p, n, k, s.data_ptr(), _cabi.stream_ptr()))
out['row_select_mad'] = {'ms': ms}
ys = torch.empty((p, n), dtype=DT, device=dev); yr = torch.empty((p, n), dtype=DT, device=dev); w = torch.empty(p, dtype=DT, device=dev)
r = torch.ones(n, dtype=DT, device=dev)
ms = timed(lambda: L.lcgp_prep_standardize(ybar.data_ptr(), c.data_ptr(), s.data_ptr(), r.data_ptr(), p, n, ys.data_ptr(), yr.data_ptr(), w.data_ptr(), _cabi.stream_ptr()))
out['standardize'] = {'ms': ms, 'gbs': 3 * 8.0 * p * n / (ms * 1e-3) / 1e9}
print(json.dumps(out))
