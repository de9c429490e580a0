"""GPU parity: padded-flat tcgen05 convolutions (conv_flat.cu: forward / data gradient / weight gradient) vs fp32 F.conv2d autograd.

The flat kernels address every tap as a ROW OFFSET into one shared-memory halo buffer (UMMA descriptor start address not aligned to the
1024-byte swizzle atom, base_offset = 0: the hardware swizzles on the absolute address — probed in profiles/r01/probe_flat.log).
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


def rnd(*shape, seed=0, scale=1.0):
    return torch.from_numpy(np.random.RandomState(seed).randn(*shape).astype(np.float32) * scale)


def _conv_case(B, Ci, Co, H, W, k, seed=0):
    pad = k // 2
    x = rnd(B, Ci, H, W, seed=41 + seed).requires_grad_(True)
    w = (rnd(Co, Ci, k, k, seed=42 + seed) / np.sqrt(Ci * k * k)).requires_grad_(True)
    b = rnd(Co, seed=43 + seed, scale=0.1)
    y = F.conv2d(x, w, b, padding=pad)
    gy = rnd(*y.shape, seed=44 + seed)
    (y * gy).sum().backward()
    return x, w, b, y, gy, pad


def _flat_fwd(ops, xd, w, b, Co, k, pad, lo, **kw):
    d = xd.device
    (wf, wdg) = ops.pack_conv_weights_bf16(w.detach().to(d), True, lo)
    xp = ops.split_bf16_flat(xd, pad, lo)
    return ops.conv_flat(xp, wf, None if b is None else b.to(d), Co, k, k, pad, pad, **kw), wdg


@pytest.mark.parametrize('B,Ci,Co,H,W,k', [
    (3, 64, 64, 16, 64, 3),       # SRB conv (whole 3x3 halo resident, 3 tiles per CTA)
    (2, 64, 64, 16, 64, 1),       # GruBlock 1x1 conv (P = 0: flat == dense)
    (2, 64, 192, 16, 64, 1),      # GRU input projection (two N tiles of 96)
    (2, 64, 256, 16, 64, 3),      # upsample conv (two N tiles of 128)
    (1, 64, 4, 32, 128, 9),       # final conv: N = 16, one kernel row per A stage, 9 taps per B stage
    (2, 192, 64, 16, 64, 1),      # pointwise with three channel blocks (data gradient of the GRU input projection)
    (3, 64, 64, 5, 17, 3),        # ragged map (rows / columns that are no multiple of anything)
    (5, 64, 128, 16, 50, 3),      # CRNN conv1 geometry
    (7, 512, 37, 1, 1, 1),        # linear layer as a pointwise conv (8 channel blocks, ragged N)
    (24, 64, 4, 32, 128, 9),      # final conv at a batch large enough for 2 tiles per CTA
    (6, 128, 256, 8, 25, 3),      # CRNN conv2: two channel blocks x whole 3x3 halo per A stage
    (5, 256, 256, 8, 25, 3),      # CRNN conv3
    (7, 256, 512, 4, 26, 3),      # CRNN conv4: four N tiles of 128 (stacked N = 256 MMAs)
    (9, 512, 512, 4, 26, 3),      # CRNN conv5: K = 4608, eight channel blocks
    (48, 512, 512, 4, 26, 3),     # ... at the bench batch (two tiles per CTA)
    (4, 128, 256, 2, 8, 3),       # STN head conv 128 -> 256 on a 2x8 map
    (4, 256, 256, 1, 4, 3),       # STN head conv on a 1x4 map (more pad rows than pixels)
])
@pytest.mark.parametrize('mode', ['bf16x3', 'bf16'])
def test_conv_flat_fwd_dgrad(B, Ci, Co, H, W, k, mode):
    from tpgsr_b200 import ops
    x, w, b, y, gy, pad = _conv_case(B, Ci, Co, H, W, k)
    d = torch.device('cuda:0')
    old = ops.CONV_MODE
    ops.CONV_MODE = mode
    lo = mode == 'bf16x3'
    tol = 3e-5 if lo else 8e-3
    try:
        assert ops.flat_supported(B, H, W, pad, Ci, Co, k, k, pad, pad)
        xd, gyd = nhwc(x.detach()).to(d), nhwc(gy).to(d)
        yd, wdg = _flat_fwd(ops, xd, w, b, Co, k, pad, lo)
        assert rel_err(nchw(yd), y) < tol, rel_err(nchw(yd), y)
        y2, _ = _flat_fwd(ops, xd, w, b, Co, k, pad, lo, act=ops.ACT_TANH, out_nchw=True)
        assert rel_err(y2, torch.tanh(y)) < tol
        if ops.flat_supported(B, H, W, pad, Co, Ci, k, k, k - 1 - pad, k - 1 - pad):
            gp = ops.split_bf16_flat(gyd, pad, lo)
            dx = ops.conv_flat(gp, wdg, None, Ci, k, k, k - 1 - pad, k - 1 - pad)
            assert rel_err(nchw(dx), x.grad) < tol, rel_err(nchw(dx), x.grad)
            base = torch.ones_like(dx)
            ops.conv_flat(gp, wdg, None, Ci, k, k, k - 1 - pad, k - 1 - pad, out=base, accumulate=True)
            assert rel_err(base - 1, nhwc(x.grad)) < 10 * tol
    finally:
        ops.CONV_MODE = old


@pytest.mark.parametrize('B,Co,H,W,k,Ci', [
    (3, 64, 16, 64, 3, 64),       # SRB conv
    (2, 64, 16, 64, 1, 64),       # GruBlock 1x1 conv
    (2, 192, 16, 64, 1, 64),      # GRU input projection (three N tiles of 64)
    (2, 256, 16, 64, 3, 64),      # upsample conv
    (1, 4, 32, 128, 9, 64),       # final conv: Co = 4 padded to 64, 9 kw taps in two groups
    (3, 64, 5, 17, 3, 64),        # ragged map
    (48, 64, 16, 64, 3, 64),      # full batch: 47 pixel splits
    (16, 4, 32, 128, 9, 64),      # final conv, many K blocks per split
    (6, 256, 8, 25, 3, 128),      # CRNN conv2: two input channel blocks
    (9, 512, 4, 26, 3, 512),      # CRNN conv5: eight channel blocks, four N tiles
    (48, 512, 4, 26, 3, 256),     # CRNN conv4 at the bench batch
    (4, 256, 1, 4, 3, 256),       # STN head conv on a 1x4 map
])
@pytest.mark.parametrize('mode', ['bf16x3', 'bf16'])
def test_wgrad_flat(B, Co, H, W, k, Ci, mode):
    from tpgsr_b200 import ops
    pad = k // 2
    x = rnd(B, Ci, H, W, seed=51)
    w = (rnd(Co, Ci, k, k, seed=52) / np.sqrt(Ci * k * k)).requires_grad_(True)
    y = F.conv2d(x, w, None, padding=pad)
    gy = rnd(*y.shape, seed=53)
    (y * gy).sum().backward()
    d = torch.device('cuda:0')
    old = ops.CONV_MODE
    ops.CONV_MODE = mode
    try:
        assert ops.wgrad_flat_supported(B, H, W, pad, Ci, Co, k, k)
        lo = mode == 'bf16x3'
        xp = ops.split_bf16_flat(nhwc(x).to(d), pad, lo)
        dyp = ops.split_bf16_flat(nhwc(gy).to(d), pad, lo, (Co + 63) // 64 * 64)
        dw = torch.full((Co, Ci, k, k), 3.0, device=d)
        ops.conv_wgrad_flat(xp, dyp, dw, Co, k, k, pad, pad, False)
        tol = 3e-5 if lo else 8e-3
        assert rel_err(dw, w.grad) < tol, rel_err(dw, w.grad)
        ops.conv_wgrad_flat(xp, dyp, dw, Co, k, k, pad, pad, True)
        assert rel_err(dw, 2 * w.grad) < tol
    finally:
        ops.CONV_MODE = old


def test_flat_full_batch_properties():
    """BASELINE batch (B=48): linearity in the input and agreement with the per-tap TMA kernel (conv_tc) on the same planes' source."""
    from tpgsr_b200 import ops
    d = torch.device('cuda:0')
    B, C, H, W = 48, 64, 16, 64
    g = torch.Generator(device='cpu').manual_seed(7)
    x1 = torch.randn(B, H, W, C, generator=g).to(d)
    x2 = torch.randn(B, H, W, C, generator=g).to(d)
    w = (torch.randn(C, C, 3, 3, generator=g) / 24.0).to(d)
    (wf, _) = ops.pack_conv_weights_bf16(w, False, True)
    f = lambda t: ops.conv_flat(ops.split_bf16_flat(t, 1, True), wf, None, C, 3, 3, 1, 1)
    y1, y2, y12 = f(x1), f(x2), f(x1 + x2)
    assert rel_err(y1 + y2, y12) < 2e-5
    y_tc = ops.conv_tc(ops.split_bf16(x1, True), wf, None, B, H, W, C, C, 3, 3, 1, 1)
    assert rel_err(y1, y_tc) < 6e-6          # conv_tc: three products; the weights-stationary plan adds the fourth (W_lo X_lo)
    ref = nhwc(F.conv2d(nchw(x1).cpu(), w.cpu(), None, padding=1))
    assert rel_err(y1, ref) < 3e-5


@pytest.mark.parametrize('B,Ci,Co,H,W,k,nchw_out', [
    (2, 64, 4, 32, 128, 9, True),     # final conv: narrow output, tanh, NCHW result
    (24, 64, 4, 32, 128, 9, True),    # ... at a batch with 3 tiles per CTA
    (3, 4, 64, 16, 64, 9, False),     # block1: narrow input (+ data gradient for the STN branch = narrow output)
    (2, 3, 64, 8, 24, 5, False),      # 5x5 with 3 input channels, ragged map
])
@pytest.mark.parametrize('mode', ['bf16x3', 'bf16'])
def test_kwstack_conv_through_engine(B, Ci, Co, H, W, k, nchw_out, mode):
    """9x9 / 5x5 convs with a narrow side through engine.conv_fwd / conv_bwd (kw-stacking -> 9x1 tcgen05 convs) vs fp32 autograd."""
    from tpgsr_b200 import engine, ops
    pad = k // 2
    x = rnd(B, Ci, H, W, seed=61).requires_grad_(True)
    w = (rnd(Co, Ci, k, k, seed=62) / np.sqrt(Ci * k * k)).requires_grad_(True)
    b = rnd(Co, seed=63, scale=0.1).requires_grad_(True)
    y = torch.tanh(F.conv2d(x, w, b, padding=pad)) if nchw_out else F.conv2d(x, w, b, padding=pad)
    gy = rnd(*y.shape, seed=64)
    (y * gy).sum().backward()
    d = torch.device('cuda:0')
    old = ops.CONV_MODE
    ops.CONV_MODE = mode
    tol = 5e-5 if mode == 'bf16x3' else 1e-2
    try:
        assert ops.kwstack_kind(B, H, W, Ci, Co, k, pad) == ('N' if Ci == 64 else 'I')
        params = {'w': w.detach().to(d), 'b': b.detach().to(d)}
        grads = {'w': torch.full_like(params['w'], 7.0), 'b': torch.zeros_like(params['b'])}
        ctx = engine.Ctx(params, grads, True)
        xd = nhwc(x.detach()).to(d)
        if nchw_out:
            yd, saved = engine.conv_fwd(ctx, xd, 'w', 'b', k, pad, act=ops.ACT_TANH, out_nchw=True)
            assert rel_err(yd, y) < tol, rel_err(yd, y)
            dpre = nhwc(gy * (1 - y.detach() ** 2)).to(d)          # gradient w.r.t. the pre-activation
        else:
            yd, saved = engine.conv_fwd(ctx, xd, 'w', 'b', k, pad)
            assert rel_err(nchw(yd), y) < tol, rel_err(nchw(yd), y)
            dpre = nhwc(gy).to(d)
        dx = engine.conv_bwd(ctx, saved, dpre, need_dx=True)
        assert rel_err(grads['w'], w.grad) < tol, rel_err(grads['w'], w.grad)
        assert rel_err(grads['b'], b.grad) < 1e-4
        assert rel_err(nchw(dx), x.grad) < tol, rel_err(nchw(dx), x.grad)
        # second backward accumulates into the same gradient buffers
        engine.conv_bwd(ctx, saved, dpre, need_dx=False)
        assert rel_err(grads['w'], 2 * w.grad) < tol
    finally:
        ops.CONV_MODE = old


@pytest.mark.parametrize('B,Ci,Co,H,W,k,pad', [
    (5, 512, 512, 2, 27, 2, 0),       # CRNN conv6: 'valid' 2x2 (its data gradient is a 'full' conv, its dY planes are padded up to x's grid)
    (48, 512, 512, 2, 27, 2, 0),      # ... at the bench batch
    (3, 64, 64, 6, 9, 3, 0),          # 'valid' 3x3
    (3, 128, 64, 1, 51, 3, 1),        # a 1-row map ('same' 1x3 in effect: InfoGen's equivalent convs use (1,3) kernels, see test_gpu_prior)
])
@pytest.mark.parametrize('mode', ['bf16x3', 'bf16'])
def test_valid_and_full_convs_through_engine(B, Ci, Co, H, W, k, pad, mode):
    """Non-'same' stride-1 convs on the flat tcgen05 kernels through engine.conv_fwd / conv_bwd (forward, dX, dW, db) vs fp32 autograd."""
    from tpgsr_b200 import engine, ops
    x = rnd(B, Ci, H, W, seed=71).requires_grad_(True)
    w = (rnd(Co, Ci, k, k, seed=72) / np.sqrt(Ci * k * k)).requires_grad_(True)
    b = rnd(Co, seed=73, scale=0.1).requires_grad_(True)
    y = F.conv2d(x, w, b, padding=pad)
    gy = rnd(*y.shape, seed=74)
    (y * gy).sum().backward()
    d = torch.device('cuda:0')
    old = ops.CONV_MODE
    ops.CONV_MODE = mode
    tol = 5e-5 if mode == 'bf16x3' else 1e-2
    try:
        params = {'w': w.detach().to(d), 'b': b.detach().to(d)}
        grads = {'w': torch.full_like(params['w'], 7.0), 'b': torch.zeros_like(params['b'])}
        ctx = engine.Ctx(params, grads, True)
        xd = nhwc(x.detach()).to(d)
        assert engine._flat_ok(xd, Ci, Co, k, k, pad, pad, 0)
        l0 = ops.LAUNCHES
        yd, saved = engine.conv_fwd(ctx, xd, 'w', 'b', k, pad)
        assert yd.shape == (B, y.shape[2], y.shape[3], Co)
        assert rel_err(nchw(yd), y) < tol, rel_err(nchw(yd), y)
        dx = engine.conv_bwd(ctx, saved, nhwc(gy).to(d), need_dx=True)
        ctx.join()
        assert rel_err(grads['w'], w.grad) < tol, rel_err(grads['w'], w.grad)
        assert rel_err(grads['b'], b.grad) < 1e-4
        assert rel_err(nchw(dx), x.grad) < tol, rel_err(nchw(dx), x.grad)
    finally:
        ops.CONV_MODE = old


@pytest.mark.parametrize('B,Co,H,W,KH', [
    (2, 36, 32, 128, 9),      # final conv after kw-stacking: 9 x 1, N = 36 partial sums
    (48, 36, 32, 128, 9),     # ... at the bench batch
    (3, 64, 16, 64, 9),       # block1 after kw-stacking (whole image height in one tile row)
    (2, 64, 11, 29, 5),       # ragged map: rows / columns that are no multiple of the 16 x 8 tile
    (1, 64, 40, 20, 7),
])
@pytest.mark.parametrize('mode', ['bf16x3', 'bf16'])
def test_tall_plan_kh_by_1_convs(B, Co, H, W, KH, mode):
    """KH x 1 convs on 64 channels run the TALL plan (2-D halo tile through one 4-D TMA box, A descriptor SBO = tile pitch): against fp32
    F.conv2d, and bit for bit against the flat plan (debug bit 3 disables the tall plan) — same MMAs in the same order, different tiling."""
    from tpgsr_b200 import ops
    from tpgsr_b200._lib import LIB
    import ctypes
    d = torch.device('cuda:0')
    ph = KH // 2
    x = rnd(B, 64, H, W, seed=91)
    w = rnd(Co, 64, KH, 1, seed=92) / np.sqrt(64 * KH)
    b = rnd(Co, seed=93, scale=0.1)
    want = F.conv2d(x, w, b, padding=(ph, 0))
    old = ops.CONV_MODE
    ops.CONV_MODE = mode
    lo = mode == 'bf16x3'
    try:
        out = (ctypes.c_int * 13)()
        assert LIB.tpgsr_conv2d_flat_plan(B, H, W, ph, 64, Co, KH, 1, ph, 0, 3 if lo else 1, out) > 0 and out[12] == 2      # the tall plan is chosen
        (wf, _) = ops.pack_conv_weights_bf16(w.to(d), False, lo)
        xp = ops.split_bf16_flat(nhwc(x).to(d), ph, lo)
        y = ops.conv_flat(xp, wf, b.to(d), Co, KH, 1, ph, 0)
        assert rel_err(nchw(y), want) < (3e-5 if lo else 8e-3), rel_err(nchw(y), want)
        y_nchw = ops.conv_flat(xp, wf, b.to(d), Co, KH, 1, ph, 0, act=ops.ACT_TANH, out_nchw=True)
        assert rel_err(y_nchw, torch.tanh(want)) < (3e-5 if lo else 8e-3)
        try:
            LIB.tpgsr_conv2d_flat_set_debug_skip(8)
            assert LIB.tpgsr_conv2d_flat_plan(B, H, W, ph, 64, Co, KH, 1, ph, 0, 3 if lo else 1, out) > 0 and out[12] != 2
            y_flat = ops.conv_flat(xp, wf, b.to(d), Co, KH, 1, ph, 0)
        finally:
            LIB.tpgsr_conv2d_flat_set_debug_skip(0)
        assert torch.equal(y, y_flat)
    finally:
        ops.CONV_MODE = old


@pytest.mark.parametrize('B,H,W,KH,KW,ph,pw', [
    (3, 16, 64, 3, 3, 1, 1),      # SRB conv, a few CTAs with 1-2 tiles
    (48, 16, 64, 3, 3, 1, 1),     # the bench shape: 139 CTAs x 4 tiles of 96 pixels
    (200, 16, 64, 3, 3, 1, 1),    # several waves of CTAs
    (3, 5, 17, 3, 3, 1, 1),       # ragged map
    (2, 32, 128, 3, 3, 1, 1),     # wide rows: the halo of a tile spans 260 extra rows
    (3, 6, 9, 3, 3, 0, 0),        # 'valid'
    (3, 6, 9, 3, 3, 2, 2),        # 'full' (the data gradient of a valid conv)
    (4, 1, 51, 1, 3, 0, 1),       # 1x3 on a one-row map
    (4, 9, 33, 3, 1, 1, 0),       # 3x1
    (2, 8, 24, 2, 2, 0, 0),       # 2x2 valid
    (2, 8, 24, 3, 4, 1, 2),       # 12 taps: the most that fit beside two 64-pixel accumulators
    (48, 16, 64, 1, 1, 0, 0),     # GruBlock 1x1 conv at the bench batch: no halo, one tap
    (3, 5, 17, 1, 1, 0, 0),       # ... on a ragged map
])
def test_weights_stationary_plan(B, H, W, KH, KW, ph, pw):
    """64 -> 64 convs with 3..12 taps run the weights-stationary plan (split weights resident in tensor memory as the A operand, pixels
    as the N side, double-buffered accumulators): against fp32 F.conv2d, against the flat plan on the same planes (debug bit 4 disables
    the plan; the flat plan omits the W_lo X_lo product, so the two agree to ~3e-6, not bit for bit), with every epilogue option."""
    from tpgsr_b200 import ops
    from tpgsr_b200._lib import LIB
    import ctypes
    d = torch.device('cuda:0')
    P = max(ph, pw)
    x = rnd(B, 64, H, W, seed=101)
    w = rnd(64, 64, KH, KW, seed=102) / np.sqrt(64 * KH * KW)
    b = rnd(64, seed=103, scale=0.1)
    want = nhwc(F.conv2d(x, w, b, padding=(ph, pw)))
    assert ops.CONV_MODE == 'bf16x3'
    out = (ctypes.c_int * 13)()
    assert LIB.tpgsr_conv2d_flat_plan(B, H, W, P, 64, 64, KH, KW, ph, pw, 3, out) > 0 and out[12] == 3      # the plan is chosen
    assert out[3] in (64, 96) and KH * KW * 32 + 2 * out[3] <= 512
    (wf, _) = ops.pack_conv_weights_bf16(w.to(d), False, True)
    xp = ops.split_bf16_flat(nhwc(x).to(d), P, True)
    bd = b.to(d)
    y = ops.conv_flat(xp, wf, bd, 64, KH, KW, ph, pw)
    assert rel_err(y, want) < 3e-5, rel_err(y, want)
    try:
        LIB.tpgsr_conv2d_flat_set_debug_skip(16)
        assert LIB.tpgsr_conv2d_flat_plan(B, H, W, P, 64, 64, KH, KW, ph, pw, 3, out) > 0 and out[12] != 3
        y_flat = ops.conv_flat(xp, wf, bd, 64, KH, KW, ph, pw)
    finally:
        LIB.tpgsr_conv2d_flat_set_debug_skip(0)
    assert rel_err(y, y_flat) < 6e-6, rel_err(y, y_flat)
    # activation, accumulate into a pitched buffer, no bias
    y_m = ops.conv_flat(xp, wf, bd, 64, KH, KW, ph, pw, act=ops.ACT_MISH)
    assert rel_err(y_m, F.mish(want)) < 3e-5
    base = torch.ones_like(y)
    ops.conv_flat(xp, wf, None, 64, KH, KW, ph, pw, out=base, accumulate=True)
    assert rel_err(base - 1, want - b.view(1, 1, 1, -1)) < 3e-4
    if (ph, pw) == ((KH - 1) // 2, (KW - 1) // 2) and KH % 2 and KW % 2:
        # the output's own bf16 hi/lo planes, written by the epilogue
        y_r, pl = ops.conv_flat(xp, wf, bd, 64, KH, KW, ph, pw, act=ops.ACT_RELU, emit_planes=True)
        assert torch.equal(y_r, torch.relu(y))
        assert torch.equal(pl.hi.float().view_as(y_r), y_r.bfloat16().float())
        assert rel_err(pl.hi.float().view_as(y_r) + pl.lo.float().view_as(y_r), y_r) < 1e-5
