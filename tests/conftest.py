"""pytest config: `gpu` marker, golden-fixture helpers, shared synthetic-weight builders."""
import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for it in items:
        if 'gpu' in it.keywords:
            it.add_marker(skip)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope='session')
def shapes():
    with open(os.path.join(GOLDEN, 'shapes.json')) as f:
        return json.load(f)


def fingerprint(t):
    f = t.detach().double().reshape(-1).cpu()
    head = torch.zeros(8, dtype=torch.float64)
    head[: min(8, f.numel())] = f[:8]
    return torch.cat([torch.stack([f.norm(), f.sum()]), head]).numpy()


def assert_fp_close(got, want, name, rtol=2e-3, atol=2e-5, noise_floor=2e-3, check_head=True):
    """Compare a 10-number fingerprint (norm, sum, first 8) with tolerances scaled by the tensor norm."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    if abs(want[0]) < noise_floor:
        # mathematically-zero gradients (e.g. a conv bias feeding a train-mode BatchNorm): only round-off noise
        assert abs(got[0]) < 20 * noise_floor, f'{name}: expected ~0, norm {got[0]} (ref {want[0]})'
        return
    scale = max(abs(want[0]), 1e-12)
    assert abs(got[0] - want[0]) <= rtol * scale + atol, f'{name}: norm {got[0]} vs {want[0]}'
    # the sum is a cancellation-prone statistic: tolerate rtol*norm*few
    if not check_head:          # norm-only mode (chaotic trajectories: element signs of +-lr Adam steps differ)
        return
    assert abs(got[1] - want[1]) <= 20 * rtol * scale + atol, f'{name}: sum {got[1]} vs {want[1]}'
    assert np.all(np.abs(got[2:] - want[2:]) <= rtol * np.maximum(np.abs(want[2:]), scale * 5e-2) + atol), \
        f'{name}: head {got[2:]} vs {want[2:]}'


def rel_err(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))
