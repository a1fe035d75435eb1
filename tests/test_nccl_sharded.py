"""Real-NCCL parity of the sharded evaluation path: sharded == unsharded == oracle, lock-step fit, and a collective
Cholesky-failure report (tests/nccl_worker.py, one process per GPU).  Needs >= 2 visible GPUs; skipped on one."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(900)
def test_sharded_equals_unsharded_equals_oracle_on_nccl():
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip('needs at least 2 GPUs (one NCCL rank per GPU)')
    world = 2 if ngpu < 4 else 4
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={world}',
           '--master-addr', '127.0.0.1', '--master-port', str(port), os.path.join(ROOT, 'tests', 'nccl_worker.py')]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=850, cwd=ROOT)
    print(res.stdout[-4000:]); print(res.stderr[-2000:])
    assert res.returncode == 0 and 'NCCL_SHARDED PASS' in res.stdout
