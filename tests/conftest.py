import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu on the GPU box')


@pytest.fixture(autouse=True, scope='session')
def _torch_threads():
    import torch
    torch.set_num_threads(min(4, os.cpu_count() or 1))   # tiny matrices: more threads only add overhead
    yield
