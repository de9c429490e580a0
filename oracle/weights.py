"""Deterministic synthetic weights / inputs, keyed by state-dict name (TEST INFRASTRUCTURE).

The same generator is used (a) by ``oracle/gen_golden.py`` to fill the *reference* modules
before recording fixtures, (b) by the tests to fill the oracle and the CUDA modules, so that
fixtures only need to store outputs.  Values depend only on (name, shape, seed) — not on
module construction order or torch's RNG stream.
"""
from __future__ import annotations

import zlib
from typing import Dict, Iterable, Tuple

import numpy as np
import torch


def _rng(name: str, seed: int) -> np.random.RandomState:
    return np.random.RandomState((zlib.crc32(name.encode()) + 7919 * seed) % (2 ** 31))


def synth_param(name: str, shape: Tuple[int, ...], seed: int = 0) -> torch.Tensor:
    r = _rng(name, seed)
    leaf = name.rsplit('.', 1)[-1]
    shape = tuple(shape)
    if leaf == 'num_batches_tracked':
        return torch.zeros((), dtype=torch.int64)
    if leaf == 'running_mean':
        return torch.tensor(0.1 * r.randn(*shape), dtype=torch.float32)
    if leaf == 'running_var':
        return torch.tensor(1.0 + 0.2 * np.abs(r.randn(*shape)), dtype=torch.float32)
    if 'weight_ih' in leaf or 'weight_hh' in leaf or 'bias_ih' in leaf or 'bias_hh' in leaf:
        hid = shape[0] // (4 if 'rnn.' in name and 'gru' not in name else 3)
        k = 1.0 / np.sqrt(hid)
        return torch.tensor(r.uniform(-k, k, size=shape), dtype=torch.float32)
    if len(shape) == 1:
        if leaf == 'weight':            # BN gamma or PReLU slope
            if shape == (1,):
                return torch.tensor([0.25], dtype=torch.float32)
            return torch.tensor(1.0 + 0.1 * r.randn(*shape), dtype=torch.float32)
        return torch.tensor(0.05 * r.randn(*shape), dtype=torch.float32)      # biases / BN beta
    if name.endswith('stn_fc2.weight'):
        return torch.tensor(0.02 * r.randn(*shape), dtype=torch.float32)
    if len(shape) == 4 and 'tconv' in name:                                      # [Cin,Cout,kh,kw]
        fan = shape[0] * shape[2] * shape[3] / 4.0
        return torch.tensor(r.randn(*shape) * np.sqrt(2.0 / fan), dtype=torch.float32)
    if len(shape) == 4:
        fan = shape[1] * shape[2] * shape[3]
        return torch.tensor(r.randn(*shape) * np.sqrt(1.5 / fan), dtype=torch.float32)
    if len(shape) == 2:
        return torch.tensor(r.randn(*shape) * np.sqrt(1.0 / shape[1]), dtype=torch.float32)
    raise ValueError(f'no rule for {name} {shape}')


def stn_fc2_bias(num_ctrl: int = 20, margin: float = 0.01) -> torch.Tensor:
    """STNHead.init_stn control points (model/stn_head.py:73-90)."""
    k = num_ctrl // 2
    xs = np.linspace(margin, 1.0 - margin, k)
    top = np.stack([xs, np.ones(k) * margin], 1)
    bot = np.stack([xs, np.ones(k) * (1 - margin)], 1)
    return torch.tensor(np.concatenate([top, bot], 0).astype(np.float32)).reshape(-1)


def synth_state_dict(shapes: Dict[str, Iterable[int]], seed: int = 0, skip_prefixes=('tps.',)) -> Dict[str, torch.Tensor]:
    out = {}
    for name, shp in shapes.items():
        if any(name.startswith(s) for s in skip_prefixes):
            continue
        if name.endswith('stn_fc2.bias'):
            out[name] = stn_fc2_bias(int(np.prod(shp)) // 2) + 0.01 * synth_param(name, tuple(shp), seed) / 0.05
            continue
        out[name] = synth_param(name, tuple(shp), seed)
    return out


def synth_images(batch: int, seed: int = 1234, smooth: bool = False):
    """SURVEY §8d input contract: LR U[0,1) [B,4,16,64], HR U[0,1) [B,4,32,128], mask channel binarised."""
    r = np.random.RandomState(seed)
    lr = r.rand(batch, 4, 16, 64).astype(np.float32)
    hr = r.rand(batch, 4, 32, 128).astype(np.float32)
    if smooth:
        def blur(a):
            for ax in (2, 3):
                a = (np.roll(a, 1, ax) + 2 * a + np.roll(a, -1, ax)) / 4
            return a
        for _ in range(3):
            lr, hr = blur(lr), blur(hr)
    lr[:, 3] = (lr[:, 3] > 0.5)
    hr[:, 3] = (hr[:, 3] > 0.5)
    return torch.from_numpy(lr), torch.from_numpy(hr)
