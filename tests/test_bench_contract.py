"""bench.py's reference arm (the oracle port on the host cores) runs without a GPU: its JSON line must carry the keys
the driver reads, rank 0 alone prints, and the CUDA arm refuses to run without a device instead of falling back."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env_extra=None, timeout=600):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + args, capture_output=True, text=True,
                          env=env, timeout=timeout, cwd=ROOT)


def test_reference_arm_json_contract():
    r = _run(['--impl', 'reference', '--config', 'cfg5_one', '--steps', '1', '--warmup', '1'])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith('{')]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j['impl'] == 'reference' and j['metric'] == 'NLL+grad evals/s' and j['unit'] == 'evals/s'
    assert j['higher_is_better'] is True and j['vs_baseline'] is None and j['dtype'] == 'f64' and j['data'] == 'synthetic'
    assert j['value'] > 0 and abs(j['ms_per_eval_extrapolated'] * j['value'] - 1e3) < 1e-6 * 1e3
    assert j['steps'] == 1 and j['n_gpus'] == 1
    cb = j['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == j['value'] and 'latents' in cb['sample']
    assert j['e2e'] == {'value': j['value'], 'unit': 'evals/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert j['extrapolated'] is True and j['sampled_latents'] == 1 and j['sample_fraction'] == 1 / 8
    assert abs(j['ms_per_step'] * 8 - j['ms_per_eval_extrapolated']) < 1e-6 * j['ms_per_eval_extrapolated']   # a step = 1 of 8 latents
    cfg = j['config']
    # identical key set and values as the CUDA arm's `config` (bench.workload_desc): no arm-specific text inside
    sys.path.insert(0, ROOT)
    import bench
    assert cfg == bench.workload_desc('cfg5_one', 1024, 6, 64, 8, 'full', 1) and 'arm' in j
    assert cfg['n'] == 1024 and cfg['d'] == 6 and cfg['p'] == 64 and cfg['q'] == 8 and cfg['workload'].startswith('cfg5_one')


def test_reference_arm_is_silent_on_other_ranks():
    r = _run(['--impl', 'reference', '--config', 'cfg5_one', '--steps', '1', '--warmup', '1'], {'RANK': '1', 'WORLD_SIZE': '2'})
    assert r.returncode == 0 and r.stdout.strip() == ''


def test_cuda_arm_refuses_to_run_without_a_device():
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    r = _run(['--steps', '1', '--warmup', '1', '--no-cpu-baseline', '--no-fit'], timeout=300)
    assert r.returncode != 0
    assert 'no CPU fallback' in (r.stderr + r.stdout)
    assert not any(ln.startswith('{') for ln in r.stdout.splitlines())
