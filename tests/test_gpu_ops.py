"""GPU parity tests, op level: every C-ABI kernel family vs the CPU oracle (oracle/ops.py) on seeded inputs.
Tolerances are fp32 round-off (1e-5 .. 1e-4 relative to the tensor norm) unless stated."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device('cuda:0')


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


def rnd(*shape, seed=0, scale=1.0):
    return torch.from_numpy(np.random.RandomState(seed).randn(*shape).astype(np.float32) * scale)


@pytest.mark.parametrize('B,Ci,Co,H,W,k,pad', [
    (2, 64, 64, 16, 64, 3, 1),      # SRB conv
    (2, 4, 64, 16, 64, 9, 4),       # block1
    (1, 64, 4, 32, 128, 9, 4),      # final conv (small-N path)
    (2, 96, 64, 16, 64, 1, 0),      # gru1 1x1 conv (TL)
    (2, 64, 256, 16, 64, 3, 1),     # upsample conv
    (3, 1, 64, 32, 100, 3, 1),      # CRNN conv0 (one input channel: the dedicated 3x3 kernel)
    (2, 1, 64, 7, 13, 3, 1),        # ... on a ragged map (rows of 13 = three groups of 4 pixels + 1)
    (2, 1, 24, 9, 10, 3, 1),        # ... with 24 output channels (6 channel quads per pixel)
    (3, 512, 512, 2, 27, 2, 0),     # CRNN conv6 (valid 2x2)
    (5, 36, 40, 1, 1, 1, 0),        # linear with ragged N/K
    (2, 3, 64, 32, 128, 9, 4),      # SRCNN conv1 (Cin=3)
])
def test_conv_fwd_dgrad_wgrad(B, Ci, Co, H, W, k, pad):
    from tpgsr_b200 import ops
    x = rnd(B, Ci, H, W, seed=1).requires_grad_(True)
    w = (rnd(Co, Ci, k, k, seed=2) / np.sqrt(Ci * k * k)).requires_grad_(True)
    b = rnd(Co, seed=3, scale=0.1).requires_grad_(True)
    y = F.conv2d(x, w, b, padding=pad)
    gy = rnd(*y.shape, seed=4)
    (y * gy).sum().backward()
    d = _dev()
    xd, wd_, bd, gyd = nhwc(x.detach()).to(d), w.detach().to(d), b.detach().to(d), nhwc(gy).to(d)
    wf, wdg = ops.pack_conv_weights(wd_, True)
    yd = ops.conv_fwd(xd, wf, bd, Co, k, k, pad, pad)
    assert rel_err(nchw(yd), y) < 2e-5
    dx = ops.conv_dgrad(gyd, wdg, (H, W), Ci, k, k, pad, pad)
    assert rel_err(nchw(dx), x.grad) < 2e-5
    dw = torch.full_like(wd_, 7.0)
    ops.conv_wgrad(xd, gyd, dw, k, k, pad, pad, False)
    assert rel_err(dw, w.grad) < 5e-5
    ops.conv_wgrad(xd, gyd, dw, k, k, pad, pad, True)          # accumulate
    assert rel_err(dw, 2 * w.grad) < 5e-5
    db = torch.zeros(Co, device=d)
    ops.colsum(gyd.view(-1, Co), db, False)
    assert rel_err(db, b.grad) < 2e-5
    # NCHW epilogue + fused tanh
    y2 = ops.conv_fwd(xd, wf, bd, Co, k, k, pad, pad, act=ops.ACT_TANH, out_nchw=True)
    assert rel_err(y2, torch.tanh(y)) < 2e-5


@pytest.mark.parametrize('act', ['none', 'relu', 'mish'])
@pytest.mark.parametrize('M,C', [(2 * 16 * 64, 64), (5, 512), (777, 32)])
def test_batchnorm_train_eval(act, M, C):
    from oracle import ops as o
    from tpgsr_b200 import ops
    x = (rnd(M, C, seed=5) * 1.5 + 0.3).requires_grad_(True)
    g = (1 + 0.1 * rnd(C, seed=6)).requires_grad_(True)
    b = (0.1 * rnd(C, seed=7)).requires_grad_(True)
    rm, rv = 0.1 * rnd(C, seed=8), 1 + 0.1 * rnd(C, seed=9).abs()
    rm_o, rv_o = rm.clone(), rv.clone()
    res = rnd(M, C, seed=10)
    f = {'none': lambda t: t, 'relu': torch.relu, 'mish': o.mish}[act]
    y = f(o.batchnorm(x, g, b, rm_o, rv_o, True)) + res
    gy = rnd(M, C, seed=11)
    (y * gy).sum().backward()
    d = _dev()
    code = {'none': ops.ACT_NONE, 'relu': ops.ACT_RELU, 'mish': ops.ACT_MISH}[act]
    xd, gd, bd, rmd, rvd = x.detach().to(d), g.detach().to(d), b.detach().to(d), rm.to(d), rv.to(d)
    mean_rstd, ss = ops.bn_stats(xd, gd, bd, rmd, rvd)
    yd = ops.bn_apply(xd, ss, code, res.to(d))
    assert rel_err(yd, y) < 1e-5
    assert rel_err(rmd, rm_o) < 1e-5 and rel_err(rvd, rv_o) < 1e-5
    dg, db = torch.zeros(C, device=d), torch.zeros(C, device=d)
    dx = ops.bn_bwd(gy.to(d), xd, ss, mean_rstd, code, True, dg, db, False)
    assert rel_err(dx, x.grad) < 5e-5
    assert rel_err(dg, g.grad) < 5e-5 and rel_err(db, b.grad) < 5e-5
    # eval mode
    ye = f(o.batchnorm(x.detach(), g.detach(), b.detach(), rm_o, rv_o, False))
    mr, ss_e = ops.bn_eval_prepare(gd, bd, rmd, rvd)
    assert rel_err(ops.bn_apply(xd, ss_e, code), ye) < 1e-5


@pytest.mark.parametrize('vertical', [False, True])
def test_gru_block_recurrence(vertical):
    """BiGRU recurrence + BPTT vs the explicit oracle loop, rows (gru2) and columns (gru1, the reference's transpose)."""
    from oracle import ops as o
    from tpgsr_b200 import ops
    B, H, W, Hd = 2, 16, 64, 32
    r = np.random.RandomState(3)
    kk = 1 / np.sqrt(Hd)
    p = {}
    for suf in ('', '_reverse'):
        p['weight_ih_l0' + suf] = torch.from_numpy(r.uniform(-kk, kk, (96, 64)).astype(np.float32)).requires_grad_(True)
        p['weight_hh_l0' + suf] = torch.from_numpy(r.uniform(-kk, kk, (96, 32)).astype(np.float32)).requires_grad_(True)
        p['bias_ih_l0' + suf] = torch.from_numpy(r.uniform(-kk, kk, 96).astype(np.float32)).requires_grad_(True)
        p['bias_hh_l0' + suf] = torch.from_numpy(r.uniform(-kk, kk, 96).astype(np.float32)).requires_grad_(True)
    x = rnd(B, 64, H, W, seed=12).requires_grad_(True)
    xs = x.transpose(-1, -2) if vertical else x
    b_, c_, h_, w_ = xs.shape
    seq = xs.permute(0, 2, 3, 1).reshape(b_ * h_, w_, c_)
    y = o.bigru(seq, p, '').reshape(b_, h_, w_, 64).permute(0, 3, 1, 2)
    if vertical:
        y = y.transpose(-1, -2)
    gy = rnd(B, 64, H, W, seed=13)
    (y * gy).sum().backward()
    d = _dev()
    xd = nhwc(x.detach()).to(d)
    wih = torch.cat([p['weight_ih_l0'], p['weight_ih_l0_reverse']], 0).detach().to(d)
    bih = torch.cat([p['bias_ih_l0'], p['bias_ih_l0_reverse']], 0).detach().to(d)
    whh = torch.stack([p['weight_hh_l0'], p['weight_hh_l0_reverse']], 0).detach().to(d).contiguous()
    bhh = torch.stack([p['bias_hh_l0'], p['bias_hh_l0_reverse']], 0).detach().to(d).contiguous()
    wf, wdg = ops.pack_conv_weights(wih.view(192, 64, 1, 1), True)
    gx = ops.conv_fwd(xd, wf, bih, 192, 1, 1, 0, 0)
    res = rnd(B, H, W, 64, seed=14).to(d)
    yd, ysum, stash = ops.gru_fwd(gx.view(-1, 192), whh, bhh, B, H, W, vertical, residual=res.view(-1, 64))
    assert rel_err(nchw(yd.view(B, H, W, 64)), y) < 2e-5
    assert rel_err(ysum, yd + res.view(-1, 64)) < 1e-6
    dwhh, dbhh, dbih = torch.zeros_like(whh), torch.zeros_like(bhh), torch.zeros(192, device=d)
    dgx = ops.gru_bwd(nhwc(gy).to(d).view(-1, 64), yd, stash, whh, dwhh, dbhh, dbih, False, B, H, W, vertical)
    dx = ops.conv_dgrad(dgx.view(B, H, W, 192), wdg, (H, W), 64, 1, 1, 0, 0)
    assert rel_err(nchw(dx), x.grad) < 5e-5
    ref_dwhh = torch.stack([p['weight_hh_l0'].grad, p['weight_hh_l0_reverse'].grad], 0)
    ref_dbhh = torch.stack([p['bias_hh_l0'].grad, p['bias_hh_l0_reverse'].grad], 0)
    ref_dbih = torch.cat([p['bias_ih_l0'].grad, p['bias_ih_l0_reverse'].grad], 0)
    assert rel_err(dwhh, ref_dwhh) < 5e-5 and rel_err(dbhh, ref_dbhh) < 5e-5 and rel_err(dbih, ref_dbih) < 5e-5
    dwih = torch.zeros(192, 64, 1, 1, device=d)
    ops.conv_wgrad(xd, dgx.view(B, H, W, 192), dwih, 1, 1, 0, 0, False)
    ref_dwih = torch.cat([p['weight_ih_l0'].grad, p['weight_ih_l0_reverse'].grad], 0)
    assert rel_err(dwih.view(192, 64), ref_dwih) < 5e-5


@pytest.mark.parametrize('k,s,p,shape', [((2, 2), (2, 2), (0, 0), (2, 16, 64, 32)), ((1, 2), (1, 2), (0, 0), (2, 1, 4, 256)),
                                         ((2, 2), (2, 1), (0, 1), (3, 8, 25, 256))])
def test_maxpool(k, s, p, shape):
    from tpgsr_b200 import ops
    x = rnd(*shape, seed=15)
    xc = nchw(x).requires_grad_(True)
    y = F.max_pool2d(xc, k, s, p)
    gy = rnd(*y.shape, seed=16)
    (y * gy).sum().backward()
    d = _dev()
    yd, arg = ops.maxpool_fwd(x.to(d), k, s, p)
    assert rel_err(nchw(yd), y) == 0.0
    dx = ops.maxpool_bwd(nhwc(gy).to(d), arg, tuple(shape), k, s, p)
    assert rel_err(nchw(dx), xc.grad) < 1e-6


def test_pixelshuffle_mish_and_acts():
    from oracle import ops as o
    from tpgsr_b200 import ops
    x = rnd(2, 256, 16, 64, seed=17).requires_grad_(True)
    y = o.mish(o.pixel_shuffle(x, 2))
    gy = rnd(*y.shape, seed=18)
    (y * gy).sum().backward()
    d = _dev()
    xd = nhwc(x.detach()).to(d)
    yd = ops.pixelshuffle_mish_fwd(xd)
    assert rel_err(nchw(yd), y) < 3e-6          # mish through ex2.approx + fast division: a few ulp per element
    dx = ops.pixelshuffle_mish_bwd(nhwc(gy).to(d), xd)
    assert rel_err(nchw(dx), x.grad) < 1e-5
    # PReLU fwd/bwd with slope gradient
    a = torch.tensor([0.25], requires_grad=True)
    z = rnd(3, 16, 64, 64, seed=19).requires_grad_(True)
    yz = o.prelu(z, a)
    gz = rnd(*z.shape, seed=20)
    (yz * gz).sum().backward()
    ad = a.detach().to(d)
    assert rel_err(ops.act_fwd(z.detach().to(d), ops.ACT_PRELU, ad), yz) < 1e-6
    da = torch.zeros(1, device=d)
    dz = ops.act_bwd(gz.to(d), z.detach().to(d), ops.ACT_PRELU, ad, da, False)
    assert rel_err(dz, z.grad) < 1e-6 and rel_err(da, a.grad) < 1e-4
    # layout round trip on a strided channel-sliced view (callers pass images[:, :3])
    img = rnd(3, 4, 16, 64, seed=21).to(d)
    v = img[:, :3]
    assert torch.equal(ops.nchw_to_nhwc(v), v.permute(0, 2, 3, 1).contiguous())
    assert torch.equal(ops.nhwc_to_nchw(ops.nchw_to_nhwc(img)), img)


def test_tps_sampler_against_golden_and_oracle():
    from oracle import ops as o
    from tpgsr_b200 import ops
    g = load_golden('tps')
    d = _dev()
    inv = torch.from_numpy(g['tps.inverse_kernel']).to(d)
    rep = torch.from_numpy(g['tps.target_coordinate_repr'])
    reprT = rep.t().contiguous().to(d)
    img = torch.from_numpy(g['img'])
    ctrl = torch.from_numpy(g['ctrl'])
    out, src = ops.tps_sample_fwd(nhwc(img).to(d), ctrl.to(d), inv, reprT)
    assert rel_err(src, g['src']) < 2e-5
    assert rel_err(nchw(out), g["out"]) < 3e-4      # white-noise image: sub-pixel coordinate round-off is amplified
    dctrl = ops.tps_sample_bwd(nhwc(torch.from_numpy(g['w'])).to(d), nhwc(img).to(d), src, inv, reprT)
    # White-noise image: d(out)/d(grid) is piecewise constant per pixel cell, and `src` itself is only good to ~1e-5
    # in fp32 (ill-conditioned TPS system), so ONE sample that lands in the neighbouring cell shifts every control
    # point's gradient by ~1% (measured: feeding the reference's own `src` into this kernel gives 9e-5).
    assert rel_err(dctrl, g['grad_ctrl']) < 3e-2
    dctrl_ref_src = ops.tps_sample_bwd(nhwc(torch.from_numpy(g['w'])).to(d), nhwc(img).to(d),
                                       torch.from_numpy(g['src']).to(d), inv, reprT)
    assert rel_err(dctrl_ref_src, g['grad_ctrl']) < 5e-4
    # identity control points reproduce the pixel grid (KAT, SURVEY §4 iv)
    tcp = torch.from_numpy(g['tps.target_control_points']).expand(3, 20, 2).contiguous()
    ident, _ = ops.tps_sample_fwd(nhwc(img).to(d), tcp.to(d), inv, reprT)
    assert rel_err(nchw(ident), g['out_identity']) < 1e-4
    # smooth image: gradient w.r.t. control points is well conditioned -> tight check against the oracle
    from oracle.weights import synth_images
    lr, _ = synth_images(4, seed=2, smooth=True)
    c2 = (torch.from_numpy(g['ctrl'][:1]).expand(4, 20, 2) * 0.97 + 0.01).clone().requires_grad_(True)
    buf = {k[4:]: torch.from_numpy(v) for k, v in g.items() if k.startswith('tps.')}
    oo, _ = o.tps_forward(lr, c2, buf, 16, 64)
    w = rnd(*oo.shape, seed=22)
    (oo * w).sum().backward()
    out2, src2 = ops.tps_sample_fwd(nhwc(lr).to(d), c2.detach().to(d), inv, reprT)
    assert rel_err(nchw(out2), oo) < 1e-4          # `src` is good to ~1e-5 in fp32 (ill-conditioned TPS solve)
    dc2 = ops.tps_sample_bwd(nhwc(w).to(d), nhwc(lr).to(d), src2, inv, reprT)
    assert rel_err(dc2, c2.grad) < 2e-4


def test_losses_against_golden():
    from tpgsr_b200 import ops
    d = _dev()
    g = load_golden('image_loss')
    sr, hr = torch.from_numpy(g['sr']).to(d), torch.from_numpy(g['hr']).to(d)
    loss = ops.image_loss_fwd(sr, hr, True, 1.0, 1e-4)
    assert abs(float(loss) - float(g['loss'])) < 2e-6 * abs(float(g['loss']))
    dsr = ops.image_loss_bwd(sr, hr, True, 1.0, 1e-4)
    assert rel_err(dsr, g['grad']) < 2e-5
    assert abs(float(ops.image_loss_fwd(sr, hr, False, 1.0, 1e-4)) - float(g['loss_nograd'])) < 1e-7
    # x100 scale + accumulate + upstream
    l2 = ops.image_loss_fwd(sr, hr, True, 1.0, 1e-4, out_scale=100.0, loss=loss, accumulate=True)
    assert abs(float(l2) - 101 * float(g['loss'])) < 1e-4 * 101 * float(g['loss'])
    up = torch.tensor([2.0], device=d)
    assert rel_err(ops.image_loss_bwd(sr, hr, True, 1.0, 1e-4, 100.0, upstream=up), 200 * torch.from_numpy(g['grad'])) < 2e-5
    s = load_golden('semantic_loss')
    pr, gt = torch.from_numpy(s['pred']).to(d), torch.from_numpy(s['gt']).to(d)
    assert abs(float(ops.semantic_loss_fwd(pr, gt)) - float(s['loss'])) < 2e-6
    assert rel_err(ops.semantic_loss_bwd(pr, gt), s['grad']) < 2e-5


def test_clip_and_adam_against_oracle():
    from oracle import ops as o
    from tpgsr_b200.optim import FlatAdam
    d = _dev()
    n, n_clip = 10007, 6001
    p = rnd(n, seed=30)
    opt = FlatAdam(p.clone().to(d), n_clipped=n_clip, lr=1e-3, betas=(0.5, 0.999), max_norm=0.25)
    p_o, m_o, v_o = p.clone(), torch.zeros(n), torch.zeros(n)
    for step in range(1, 4):
        g = rnd(n, seed=30 + step, scale=0.01 * step)
        opt.grads.copy_(g.to(d))
        opt.step()
        g_o = g.clone()
        norm = o.clip_grad_norm([g_o[:n_clip]], 0.25)
        o.adam_step(p_o, g_o, m_o, v_o, step)
        assert abs(float(opt.state[3]) - float(norm)) < 1e-5 * float(norm)
        assert rel_err(opt.params, p_o) < 1e-6


@pytest.mark.parametrize('B,Ci,Co,H,W,k', [
    (3, 64, 64, 16, 64, 3),       # SRB conv
    (2, 64, 256, 16, 64, 3),      # upsample conv (N = 256)
    (2, 64, 192, 16, 64, 1),      # GRU input projection (N = 192)
    (1, 64, 4, 32, 128, 9),       # final conv: Co = 4 padded to N = 16 by TMA zero fill
    (2, 128, 64, 16, 64, 3),      # two channel blocks per tap
])
@pytest.mark.parametrize('mode', ['bf16x3', 'bf16'])
def test_conv_tensor_core_path(B, Ci, Co, H, W, k, mode):
    """tcgen05 implicit-GEMM conv (forward and data gradient) vs fp32 F.conv2d."""
    from tpgsr_b200 import ops
    pad = k // 2
    x = rnd(B, Ci, H, W, seed=41).requires_grad_(True)
    w = (rnd(Co, Ci, k, k, seed=42) / np.sqrt(Ci * k * k)).requires_grad_(True)
    b = rnd(Co, seed=43, scale=0.1)
    y = F.conv2d(x, w, b, padding=pad)
    gy = rnd(*y.shape, seed=44)
    (y * gy).sum().backward()
    d = _dev()
    old = ops.CONV_MODE
    ops.CONV_MODE = mode
    try:
        assert ops.tc_supported(B, H, W, Ci, H, W, Co, k, k)
        xd, gyd = nhwc(x.detach()).to(d), nhwc(gy).to(d)
        (wf, wdg) = ops.pack_conv_weights_bf16(w.detach().to(d), True, mode == 'bf16x3')
        yd = ops.conv_tc(ops.split_bf16(xd, mode == 'bf16x3'), wf, b.to(d), B, H, W, Ci, Co, k, k, pad, pad)
        tol = 3e-5 if mode == 'bf16x3' else 8e-3
        assert rel_err(nchw(yd), y) < tol, rel_err(nchw(yd), y)
        y2 = ops.conv_tc(ops.split_bf16(xd, mode == 'bf16x3'), wf, b.to(d), B, H, W, Ci, Co, k, k, pad, pad, act=ops.ACT_TANH, out_nchw=True)
        assert rel_err(y2, torch.tanh(y)) < tol
        if Co % 64 == 0:
            dx = ops.conv_tc(ops.split_bf16(gyd, mode == 'bf16x3'), wdg, None, B, H, W, Co, Ci, k, k, k - 1 - pad, k - 1 - pad)
            assert rel_err(nchw(dx), x.grad) < tol, rel_err(nchw(dx), x.grad)
            base = torch.ones_like(dx)
            ops.conv_tc(ops.split_bf16(gyd, mode == 'bf16x3'), wdg, None, B, H, W, Co, Ci, k, k, k - 1 - pad, k - 1 - pad, out=base, accumulate=True)
            assert rel_err(base - 1, nhwc(x.grad)) < 10 * tol
    finally:
        ops.CONV_MODE = old


def test_psnr_ssim_against_golden_and_oracle():
    """utils/ssim_psnr.py parity metrics: the fixture recorded from the reference and the oracle on a second, sliced pair."""
    from oracle import ops as o
    from tpgsr_b200.utils.ssim_psnr import SSIM, calculate_psnr
    g = load_golden('psnr_ssim')
    d = _dev()
    a, b = torch.from_numpy(g['a']).to(d), torch.from_numpy(g['b']).to(d)
    assert abs(float(calculate_psnr(a, b)) - float(g['psnr'])) < 1e-3
    assert abs(float(SSIM()(a, b)) - float(g['ssim'])) < 1e-5
    x = rnd(5, 4, 32, 128, seed=71).abs().clamp(0, 1)
    y = (x + 0.05 * rnd(5, 4, 32, 128, seed=72)).clamp(0, 1)
    assert abs(float(calculate_psnr(x.to(d), y.to(d))) - float(o.calculate_psnr(x, y))) < 1e-3
    assert abs(float(SSIM()(x.to(d), y.to(d))) - float(o.ssim(x, y))) < 1e-5
    assert float(calculate_psnr(a, a)) == float('inf')


@pytest.mark.parametrize('act', ['none', 'relu', 'mish'])
@pytest.mark.parametrize('P', [0, 1])
def test_bn_fused_passes_match_three_launch_passes_bitwise(act, P):
    """tpgsr_bn_stats_fused / tpgsr_bn_apply_planes / tpgsr_bn_bwd_fused ("last block finalizes", operand planes from the apply pass) against the
    round-1 three-launch passes followed by tpgsr_split_bf16_flat: identical bits, twice in a row (the ticket resets itself)."""
    from tpgsr_b200 import ops
    d = _dev()
    B, H, W, C = 5, 6, 9, 64
    code = {'none': ops.ACT_NONE, 'relu': ops.ACT_RELU, 'mish': ops.ACT_MISH}[act]
    x = rnd(B, H, W, C, seed=31).to(d) * 1.7 + 0.3
    res = rnd(B, H, W, C, seed=32).to(d)
    gy = rnd(B, H, W, C, seed=33).to(d)
    g, b = (1 + 0.1 * rnd(C, seed=34)).to(d), (0.1 * rnd(C, seed=35)).to(d)
    for rep in range(2):
        rm1, rv1, rm2, rv2 = torch.zeros(C, device=d), torch.ones(C, device=d), torch.zeros(C, device=d), torch.ones(C, device=d)
        old = ops.BN_FUSED
        try:
            ops.BN_FUSED = False
            mr0, ss0 = ops.bn_stats(x.view(-1, C), g, b, rm1, rv1)
            y0 = ops.bn_apply(x.view(-1, C), ss0, code, res.view(-1, C)).view(B, H, W, C)
            p0 = ops.split_bf16_flat(y0, P, True)
            dg0, db0 = torch.zeros(C, device=d), torch.zeros(C, device=d)
            dx0 = ops.bn_bwd(gy.view(-1, C), x.view(-1, C), ss0, mr0, code, True, dg0, db0, False).view(B, H, W, C)
            q0 = ops.split_bf16_flat(dx0, P, True)
            ops.BN_FUSED = rep == 0                      # rep 1: planes from the apply passes behind the stand-alone finalize kernels
            mr1, ss1 = ops.bn_stats(x.view(-1, C), g, b, rm2, rv2)
            y1, p1 = ops.bn_apply_planes(x, ss1, code, res, P)
            dg1, db1 = torch.zeros(C, device=d), torch.zeros(C, device=d)
            dx1, q1 = ops.bn_bwd_fused(gy, x, ss1, mr1, code, True, dg1, db1, False, P)
            dx2, none = ops.bn_bwd_fused(gy, x, ss1, mr1, code, True, dg1, db1, True, None)
        finally:
            ops.BN_FUSED = old
        for a_, b_ in ((mr0, mr1), (ss0, ss1), (rm1, rm2), (rv1, rv2), (y0, y1), (dx0, dx1), (dx0, dx2), (dg0 * 2, dg1), (db0 * 2, db1),
                       (p0.hi, p1.hi), (p0.lo, p1.lo), (q0.hi, q1.hi), (q0.lo, q1.lo)):
            assert torch.equal(a_, b_), (act, P, rep)
        assert none is None and (p1.B, p1.H, p1.W, p1.C, p1.P) == (B, H, W, C, P)


@pytest.mark.parametrize('C', [2, 3, 4, 7])
def test_standalone_gradient_prior_loss_covers_every_channel(C):
    """GradientPriorLoss on its own (loss/image_loss.py:33-51) averages over ALL channels of its input (ImageLoss hands it the RGB slice)."""
    from oracle import ops as o
    from tpgsr_b200.loss.image_loss import GradientPriorLoss
    a = rnd(3, C, 12, 20, seed=41).requires_grad_(True)
    b = rnd(3, C, 12, 20, seed=42)
    want = (o.gradient_map(a) - o.gradient_map(b)).abs().mean()
    want.backward()
    ad = a.detach().to(_dev()).requires_grad_(True)
    got = GradientPriorLoss()(ad, b.to(_dev()))
    got.backward()
    assert abs(float(got) - float(want)) < 2e-6 * abs(float(want)) + 1e-7
    assert rel_err(ad.grad, a.grad) < 2e-5


@pytest.mark.parametrize('B,H,W', [(3, 5, 24), (2, 16, 64), (48, 16, 64), (1, 7, 9)])
@pytest.mark.parametrize('vertical', [False, True])
def test_gru_tensor_core_forward_matches_ffma_kernel(B, H, W, vertical):
    """The tensor-core forward recurrence (mma.sync on split-bf16 h and W_hh, 16 sequences per CTA, warp-specialised I/O) against the
    warp-per-sequence FFMA kernel on the same inputs: every output (y, y + residual, the BPTT stash), ragged sequence counts, with and
    without residual / stash (inference)."""
    from tpgsr_b200 import ops
    from tpgsr_b200._lib import LIB
    d = _dev()
    M = B * H * W
    g = torch.Generator(device='cpu').manual_seed(B * 100 + H)
    gx = torch.randn(M, 192, generator=g).to(d)
    whh = (torch.rand(2, 96, 32, generator=g) * 2 - 1).mul(1 / np.sqrt(32)).to(d)
    bhh = (torch.rand(2, 96, generator=g) * 2 - 1).mul(1 / np.sqrt(32)).to(d)
    res = torch.randn(M, 64, generator=g).to(d)
    try:
        LIB.tpgsr_gru_set_mode(0)
        y0, s0, st0 = ops.gru_fwd(gx, whh, bhh, B, H, W, vertical, residual=res)
        LIB.tpgsr_gru_set_mode(2)
        y1, s1, st1 = ops.gru_fwd(gx, whh, bhh, B, H, W, vertical, residual=res)
        y2, s2, st2 = ops.gru_fwd(gx, whh, bhh, B, H, W, vertical, residual=None, need_stash=False)
    finally:
        LIB.tpgsr_gru_set_mode(1)
    assert rel_err(y1, y0) < 5e-6 and rel_err(s1, s0) < 5e-6 and rel_err(st1, st0) < 5e-6, (rel_err(y1, y0), rel_err(st1, st0))
    assert float((y1 - y0).abs().max()) < 2e-5
    assert torch.equal(y2, y1) and s2 is None and st2 is None


@pytest.mark.parametrize('B,H,W', [(3, 5, 24), (2, 16, 64), (48, 16, 64), (1, 7, 9)])
@pytest.mark.parametrize('vertical', [False, True])
def test_gru_tensor_core_bptt_matches_ffma_kernel(B, H, W, vertical):
    """The tensor-core BPTT recurrence against the warp-per-sequence FFMA kernel on the same stash: dgx and, through the shared dW_hh /
    reduction kernels, dW_hh and both bias gradients (the per-CTA partial slots of the two kernels differ; the sums must not)."""
    from tpgsr_b200 import ops
    from tpgsr_b200._lib import LIB
    d = _dev()
    M = B * H * W
    g = torch.Generator(device='cpu').manual_seed(B * 10 + W)
    gx = torch.randn(M, 192, generator=g).to(d)
    whh = (torch.rand(2, 96, 32, generator=g) * 2 - 1).mul(1 / np.sqrt(32)).to(d)
    bhh = (torch.rand(2, 96, generator=g) * 2 - 1).mul(1 / np.sqrt(32)).to(d)
    dy = torch.randn(M, 64, generator=g).to(d)
    outs = {}
    try:
        LIB.tpgsr_gru_set_mode(0)
        y, _, stash = ops.gru_fwd(gx, whh, bhh, B, H, W, vertical)
        for mode in (0, 2):
            LIB.tpgsr_gru_set_mode(mode)
            dwhh, dbhh, dbih = torch.zeros_like(whh), torch.zeros_like(bhh), torch.zeros(192, device=d)
            dgx = ops.gru_bwd(dy, y, stash, whh, dwhh, dbhh, dbih, False, B, H, W, vertical)
            dgx2, part = ops.gru_bwd_recurrence(dy, y, stash, whh, B, H, W, vertical)           # the two-call form the engine uses
            dw2, db2, di2 = torch.ones_like(whh), torch.ones_like(bhh), torch.ones(192, device=d)
            ops.gru_bwd_weights(dgx2, y, stash, part, dw2, db2, di2, True, B, H, W, vertical)
            assert torch.equal(dgx2, dgx) and rel_err(dw2 - 1, dwhh) < 1e-5 and rel_err(db2 - 1, dbhh) < 1e-5 and rel_err(di2 - 1, dbih) < 1e-5
            outs[mode] = (dgx, dwhh, dbhh, dbih)
    finally:
        LIB.tpgsr_gru_set_mode(1)
    for name, a, b in zip(('dgx', 'dwhh', 'dbhh', 'dbih'), outs[2], outs[0]):
        assert rel_err(a, b) < 1e-5, (name, rel_err(a, b))


def test_bn_statistics_of_channels_with_large_mean_to_std_ratio():
    """Train-mode BatchNorm statistics of channels whose |mean| dwarfs their standard deviation (post-ReLU maps, large conv biases): the
    kernels sum around a per-channel pivot, so E[x^2] - mean^2 does not cancel in the fp32 partial sums (ATen: Welford)."""
    from tpgsr_b200 import ops
    d = _dev()
    M, C = 48 * 16 * 64, 64
    g = torch.Generator(device='cpu').manual_seed(5)
    base = torch.randn(M, C, generator=g, dtype=torch.float64)
    mean = torch.tensor([0.0, 1.0, 30.0, 300.0, -1000.0, 3000.0] * 11)[:C].double()
    std = torch.tensor([1.0, 0.5, 0.1, 0.05, 0.02, 0.3] * 11)[:C].double()
    x64 = base * std + mean
    x = x64.float()
    want_mean, want_var = x.double().mean(0), x.double().var(0, unbiased=False)
    gamma, beta = torch.ones(C, device=d), torch.zeros(C, device=d)
    rm, rv = torch.zeros(C, device=d), torch.ones(C, device=d)
    mean_rstd, scale_shift = ops.bn_stats(x.to(d), gamma, beta, rm, rv, momentum=1.0)
    got_mean = mean_rstd[0].double().cpu()
    got_var = (1.0 / mean_rstd[1].double().cpu() ** 2) - 1e-5
    assert float(((got_mean - want_mean).abs() / (want_mean.abs() + std)).max()) < 1e-6
    assert float(((got_var - want_var).abs() / want_var).max()) < 2e-4, ((got_var - want_var).abs() / want_var).max()
    # running_var (momentum 1: the unbiased batch variance itself)
    assert float(((rv.double().cpu() - want_var * M / (M - 1)).abs() / want_var).max()) < 2e-4
