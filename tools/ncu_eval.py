"""One objective + gradient evaluation of a named configuration inside a cudaProfiler range (for ncu
--profile-from-start off):   python tools/ncu_eval.py cfg4_shard8 [warmup evaluations [q override]]   (developer tool)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lcgp_b200 import LCGP, synthetic
cfg = sys.argv[1]; warm = int(sys.argv[2]) if len(sys.argv) > 2 else 1
x, y, _, _, mk = synthetic.make_config(cfg)
if len(sys.argv) > 3: mk['q'] = int(sys.argv[3])      # latent-count override
m = LCGP(y=y, x=x, shard=False, **mk)
eng = m.engine
eng.use_plans = False                      # launch by launch, so that every kernel is visible to the profiler
lLmb, lLmb0, lsig_p, lnug = m.get_param()
for _ in range(warm):
    eng.evaluate_device(lLmb, lLmb0, lnug, lsig_p, True)
torch.cuda.synchronize()
torch.cuda.profiler.start()
out = eng.evaluate_device(lLmb, lLmb0, lnug, lsig_p, True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(cfg, 'objective', float(out[0]), flush=True)
