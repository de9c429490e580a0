"""GPU parity: tensor-core weight gradient (MN-major UMMA operands straight from NHWC bf16 planes) vs fp32 autograd."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def rnd(*shape, seed=0, scale=1.0):
    return torch.from_numpy(np.random.RandomState(seed).randn(*shape).astype(np.float32) * scale)


@pytest.mark.parametrize('B,Co,H,W,k', [
    (3, 64, 16, 64, 3),       # SRB conv
    (2, 64, 16, 64, 1),       # GruBlock 1x1 conv
    (2, 192, 16, 64, 1),      # GRU input projection (N = 192: three 64-channel groups)
    (2, 256, 16, 64, 3),      # upsample conv (two N tiles of 128)
    (1, 4, 32, 128, 9),       # final conv: Co = 4 padded to 64, 9 kw taps in two groups, half-row K blocks
])
@pytest.mark.parametrize('mode', ['bf16x3', 'bf16'])
def test_wgrad_tensor_core(B, Co, H, W, k, mode):
    from tpgsr_b200 import ops
    Ci, pad = 64, k // 2
    x = rnd(B, Ci, H, W, seed=51)
    w = (rnd(Co, Ci, k, k, seed=52) / np.sqrt(Ci * k * k)).requires_grad_(True)
    y = F.conv2d(x, w, None, padding=pad)
    gy = rnd(*y.shape, seed=53)
    (y * gy).sum().backward()
    d = torch.device('cuda:0')
    old = ops.CONV_MODE
    ops.CONV_MODE = mode
    try:
        assert ops.wgrad_tc_supported(B, H, W, Ci, Co, k, k)
        xd, gyd = nhwc(x).to(d), nhwc(gy).to(d)
        lo = mode == 'bf16x3'
        xp = ops.split_bf16(xd, lo)
        cpad = (Co + 63) // 64 * 64
        dyp = ops.split_bf16(gyd, lo) if cpad == Co else ops.split_bf16_pad(gyd, cpad, lo)
        dw = torch.full((Co, Ci, k, k), 3.0, device=d)
        ops.conv_wgrad_tc(xp, dyp, dw, B, H, W, Ci, Co, k, k, pad, pad, False)
        tol = 3e-5 if lo else 8e-3
        assert rel_err(dw, w.grad) < tol, rel_err(dw, w.grad)
        ops.conv_wgrad_tc(xp, dyp, dw, B, H, W, Ci, Co, k, k, pad, pad, True)
        assert rel_err(dw, 2 * w.grad) < tol
    finally:
        ops.CONV_MODE = old
