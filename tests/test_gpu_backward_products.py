"""GPU tests of the opt-in reduced-product backward (TPGSR_BWD_PRODUCTS=2: dY as one bf16 plane against exact weights / activations;
TPGSR_WGRAD_PRODUCTS=1: one-product weight gradient).  Default off — the parity mode keeps three products everywhere — measured on B200 in
round 2 (+6.7% / +9.3% images/s on config 3, profiles/r02/bench_c3_bwd2*_call5.json).  The kernels' nprod = 2 instantiations are exercised
here directly, and a TSRN step with the switches on is held to 1e-3 of the default step's gradients."""

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


GEOMS = [(3, 64, 64, 16, 64, 3), (2, 64, 64, 16, 64, 1), (2, 64, 192, 16, 64, 1), (2, 64, 256, 16, 64, 3), (2, 192, 64, 16, 64, 1),
         (3, 64, 64, 5, 17, 3), (48, 64, 64, 16, 64, 3), (7, 512, 37, 1, 1, 1)]


@pytest.mark.parametrize('B,Ci,Co,H,W,k', GEOMS)
def test_two_product_data_gradient(B, Ci, Co, H, W, k):
    """conv_flat with nprod = 2: A (= dY) as ONE bf16 plane against hi + lo weights == fp32 conv of the bf16-rounded dY."""
    from tpgsr_b200 import ops
    assert ops.CONV_MODE == 'bf16x3'
    d = torch.device('cuda:0')
    pad = k // 2
    x = rnd(B, Ci, H, W, seed=61)
    w = rnd(Co, Ci, k, k, seed=62) / np.sqrt(Ci * k * k)
    b = rnd(Co, seed=63, scale=0.1)
    wf, _ = ops.pack_conv_weights_bf16(w.to(d), False, True)
    xp = ops.split_bf16_flat(nhwc(x).to(d), pad, False)            # no lo plane -> the wrapper selects nprod = 2
    assert xp.lo is None and ops._nprod_for(xp.lo, wf[1]) == 2
    y = ops.conv_flat(xp, wf, b.to(d), Co, k, k, pad, pad)
    want = F.conv2d(x.bfloat16().float(), w, b, padding=pad)
    assert rel_err(y.permute(0, 3, 1, 2), want) < 3e-5, rel_err(y.permute(0, 3, 1, 2), want)
    y3 = ops.conv_flat(ops.split_bf16_flat(nhwc(x).to(d), pad, True), wf, b.to(d), Co, k, k, pad, pad)
    exact = F.conv2d(x, w, b, padding=pad)
    assert rel_err(y3.permute(0, 3, 1, 2), exact) < 3e-5 < rel_err(y.permute(0, 3, 1, 2), exact) < 4e-3      # and it IS the coarser mode


@pytest.mark.parametrize('B,Co,H,W,k', [(3, 64, 16, 64, 3), (2, 192, 16, 64, 1), (2, 256, 16, 64, 3), (1, 4, 32, 128, 9), (48, 64, 16, 64, 3)])
def test_two_product_weight_gradient(B, Co, H, W, k):
    from tpgsr_b200 import ops
    d = torch.device('cuda:0')
    Ci, pad = 64, k // 2
    x = rnd(B, Ci, H, W, seed=71)
    gy = rnd(B, Co, H, W, seed=73)
    want = torch.nn.grad.conv2d_weight(x, (Co, Ci, k, k), gy.bfloat16().float(), padding=pad)
    xp = ops.split_bf16_flat(nhwc(x).to(d), pad, True)
    dyp = ops.split_bf16_flat(nhwc(gy).to(d), pad, False, (Co + 63) // 64 * 64)
    dw = torch.full((Co, Ci, k, k), 3.0, device=d)
    ops.conv_wgrad_flat(xp, dyp, dw, Co, k, k, pad, pad, False)
    assert rel_err(dw, want) < 3e-5, rel_err(dw, want)


def test_two_product_backward_train_step_stays_close(shapes, monkeypatch):
    """TSRN train step with TPGSR_BWD_PRODUCTS=2 (and additionally TPGSR_WGRAD_PRODUCTS=1) semantics: identical loss (forward untouched),
    gradients within 1e-3 of the default step's outside the (chaotic) STN head."""
    import copy
    from oracle.weights import synth_images, synth_state_dict
    from tpgsr_b200 import ops
    from tpgsr_b200.model.tsrn import TSRN
    from tpgsr_b200.train import TSRNTrainStep
    KW = dict(scale_factor=2, width=128, height=32, STN=True, mask=True, srb_nums=5, hidden_units=32)
    d = torch.device('cuda:0')
    sd = synth_state_dict(shapes['tsrn_train'], 1)
    lr, hr = synth_images(4, seed=70)
    runs = []
    for bwd2 in (False, True, 'w1'):
        monkeypatch.setattr(ops, 'BWD2', bool(bwd2))
        monkeypatch.setattr(ops, 'WGRAD1', bwd2 == 'w1')
        net = TSRN(**KW)
        net.load_state_dict(copy.deepcopy(sd), strict=False)
        net = net.to(d).train()
        ts = TSRNTrainStep(net, batch=4, use_graph=False)
        ts.load_batch(lr.to(d), hr.to(d))
        ts._fwd_bwd()
        torch.cuda.synchronize()
        runs.append((float(ts.loss), {k: v.clone() for k, v in ts.st.gviews.items()}))
    (l0, g0) = runs[0]
    for (l1, g1) in runs[1:]:
        assert l0 == l1
        num = sum(float((g1[k] - g0[k]).norm() ** 2) for k in g0 if 'stn_head' not in k)
        den = sum(float(g0[k].norm() ** 2) for k in g0 if 'stn_head' not in k)
        assert 0 < (num / den) ** 0.5 < 1e-3, (num / den) ** 0.5
