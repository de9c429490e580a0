"""bench.py's contract on the CPU: the reference arm (the oracle port on the host cores) prints exactly one JSON line with the keys the driver
reads, on the SAME metric / workload strings as the GPU arm; the GPU arm refuses to run without a device instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(900)
def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, cwd=ROOT, timeout=850)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    sys.path.insert(0, ROOT)
    import bench
    assert d['impl'] == 'reference' and d['metric'] == bench.METRIC and d['unit'] == 'images/s' and d['higher_is_better'] is True
    cfg = bench.CONFIGS[3]                      # the default workload at every N
    assert d['config']['workload'] == cfg['workload'] and d['config']['per_gpu_batch'] == cfg['batch'] and d['config']['config_id'] == 3
    assert d['n_gpus'] == 1 and d['steps'] == 1 and d['warmup'] == 0 and d['value'] > 0 and d['ms_per_step'] > 0
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == d['value'] and 'images' in cb['sample']
    assert d['e2e'] == {'value': d['value'], 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['vs_baseline'] is None and d['gpu_launches'] == 0


def test_gpu_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '1', '--warmup', '0'], capture_output=True, text=True, cwd=ROOT,
                       timeout=300)
    assert r.returncode != 0 and 'no CPU fallback' in (r.stderr + r.stdout)
    assert not [l for l in r.stdout.splitlines() if l.startswith('{')]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, cwd=ROOT, env=env, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ''
