"""GPU-box probe: which UMMA descriptor flavour reproduces F.conv2d when the operand window starts at a row offset that is
(a) a multiple of 8 rows (1024-byte aligned: 3x1 kernel on a 63-wide map, W+P = 64) or (b) arbitrary (1x3 / 3x3 kernels)?
Prints relative errors for desc_mode 0 (base_offset = 0) and 1 (base_offset from the address), forward and weight gradient."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.nn.functional as F
from tpgsr_b200 import ops
from tpgsr_b200._lib import LIB

d = torch.device('cuda:0')
rnd = lambda *s, seed=0: torch.from_numpy(np.random.RandomState(seed).randn(*s).astype(np.float32))
nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous()
rel = lambda a, b: float((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm())

for name, (B, H, W, KH, KW) in {'3x1 aligned (W+P=64)': (2, 16, 63, 3, 1), '1x3 unaligned': (2, 16, 64, 1, 3), '3x3': (2, 16, 64, 3, 3),
                                '9x9 N=16': (1, 32, 128, 9, 9)}.items():
    Co = 4 if KH == 9 else 64
    ph, pw = KH // 2, KW // 2
    P = max(ph, pw)
    x = rnd(B, 64, H, W, seed=1)
    w = (rnd(Co, 64, KH, KW, seed=2) / np.sqrt(64 * KH * KW)).requires_grad_(True)
    y = F.conv2d(x, w, None, padding=(ph, pw))
    gy = rnd(*y.shape, seed=3)
    (y * gy).sum().backward()
    for mode in (0, 1):
        LIB.tpgsr_tc_set_desc_mode(mode)
        try:
            (wf, _) = ops.pack_conv_weights_bf16(w.detach().to(d), False, True)
            xp = ops.split_bf16_flat(nhwc(x).to(d), P, True)
            yd = ops.conv_flat(xp, wf, None, Co, KH, KW, ph, pw)
            torch.cuda.synchronize()
            e_f = rel(yd.permute(0, 3, 1, 2), y)
        except Exception as ex:      # noqa: BLE001
            e_f = repr(ex)
        try:
            dyp = ops.split_bf16_flat(nhwc(gy).to(d), P, True, (Co + 63) // 64 * 64)
            dw = torch.zeros(Co, 64, KH, KW, device=d)
            ops.conv_wgrad_flat(xp, dyp, dw, Co, KH, KW, ph, pw, False)
            torch.cuda.synchronize()
            e_w = rel(dw, w.grad)
        except Exception as ex:      # noqa: BLE001
            e_w = repr(ex)
        print(f'{name:24s} desc_mode={mode}: fwd rel err {e_f}   wgrad rel err {e_w}', flush=True)
LIB.tpgsr_tc_set_desc_mode(0)
