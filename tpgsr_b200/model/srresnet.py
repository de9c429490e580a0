"""SRResNet_TL drop-in module (reference: ``model/srresnet.py:98-234``; BASELINE config 5) executed by the sm_100a kernels.

``SRResNet_TL(scale_factor, width, height, STN, mask)`` keeps the reference's constructor, ``forward(x, text_emb=None)`` signature,
submodule names and ``state_dict`` keys.  Data flow (model/srresnet.py:139-163): conv9x9+PReLU; 5 x ResidualBlock_TL
[conv3x3+BN+PReLU -> cat(text prior, 32 ch) -> conv3x3(96->64)+BN -> +x]; conv3x3+BN, +skip; conv3x3(64->256) -> PixelShuffle(2) -> PReLU;
conv9x9(64->4); tanh.  The prior map comes from this file's InfoGen flavour (tconv1/2 WITHOUT padding -> [B,32,25,213]) and a 2-D
bilinear(align_corners=True) resize to 16x64.  The reference's STN option imports the recogniser STN head, which cannot run on a 16x64
input — like ``train_TPGSR-SRResNet.sh`` this module supports ``STN=False`` only.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
from torch import nn

from .. import engine, ops, ops_prior
from ..ops import ACT_NONE, ACT_PRELU, ACT_RELU, ACT_TANH
from ._flat import FlatState

# (stride, padding) of the four ConvTranspose2d layers, SRResNet flavour (model/srresnet.py:174-184)
_TCONV = [((2, 2), (0, 0)), ((2, 2), (0, 0)), ((2, 2), (1, 1)), ((2, 1), (1, 1))]


class ResidualBlock_TL(nn.Module):
    def __init__(self, channels, out_text_channels=32):
        super().__init__()
        self.conv1 = nn.Conv2d(channels, channels, kernel_size=3, padding=1)
        self.bn1 = nn.BatchNorm2d(channels)
        self.prelu = nn.PReLU()
        self.conv2 = nn.Conv2d(channels + out_text_channels, channels, kernel_size=3, padding=1)
        self.bn2 = nn.BatchNorm2d(channels)


class UpsampleBLock(nn.Module):
    def __init__(self, in_channels, up_scale):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, in_channels * up_scale ** 2, kernel_size=3, padding=1)
        self.pixel_shuffle = nn.PixelShuffle(up_scale)
        self.prelu = nn.PReLU()


class InfoGen(nn.Module):
    def __init__(self, t_emb, output_size):
        super().__init__()
        self.tconv1 = nn.ConvTranspose2d(t_emb, 512, 3, 2, bias=False)
        self.bn1 = nn.BatchNorm2d(512)
        self.tconv2 = nn.ConvTranspose2d(512, 128, 3, 2, bias=False)
        self.bn2 = nn.BatchNorm2d(128)
        self.tconv3 = nn.ConvTranspose2d(128, 64, 3, 2, padding=1, bias=False)
        self.bn3 = nn.BatchNorm2d(64)
        self.tconv4 = nn.ConvTranspose2d(64, output_size, 3, (2, 1), padding=1, bias=False)
        self.bn4 = nn.BatchNorm2d(output_size)


# ----------------------------------------------------------------------------------------------------------------- engine pieces
def info_gen2d_fwd(ctx: engine.Ctx, text_emb: Optional[torch.Tensor], out_hw, prior_nhwc: Optional[torch.Tensor] = None):
    """[B,37,1,26] NCHW prior -> NHWC prior map [B,H,W,32]: 4 x (ConvTranspose2d + BN + ReLU) then bilinear(align_corners=True).
    The fused train step passes the NHWC prior [B,1,26,37] the softmax / prior kernel produced (``prior_nhwc``)."""
    t = prior_nhwc if prior_nhwc is not None else ops.nchw_to_nhwc(text_emb)      # [B,1,26,37]
    saved = []
    for i, ((sh, sw), (ph, pw)) in enumerate(_TCONV, 1):
        wname = f'infoGen.tconv{i}.weight'
        wt = ctx.P(wname)
        Ci, Co, KH, KW = wt.shape
        wc = ops_prior.tconv_weight2d_fwd(wt)                        # OIHW [Co,Ci,KH,KW], flipped
        wf, wd = ops.pack_conv_weights(wc, ctx.grads is not None)
        up = ops_prior.zero_upsample2d(t, sh, sw)
        c = ops.conv_fwd(up, wf, None, Co, KH, KW, KH - 1 - ph, KW - 1 - pw)
        a, s_bn = engine.bn_fwd(ctx, c, f'infoGen.bn{i}.', ACT_RELU)
        saved.append((wname, wd, up, (t.shape[1], t.shape[2]), (sh, sw), (KH - 1 - ph, KW - 1 - pw), (KH, KW), Ci, Co, s_bn))
        t = a
    text = ops_prior.bilinear2d_fwd(t, out_hw[0], out_hw[1])
    return text, {'layers': saved, 'text': text, 'hw_last': (t.shape[1], t.shape[2])}


def info_gen2d_bwd(ctx: engine.Ctx, s_text, dtext: torch.Tensor, need_dprior: bool, nhwc_out: bool = False):
    d = ops_prior.bilinear2d_bwd(dtext, *s_text['hw_last'])
    layers = s_text['layers']
    for idx in range(len(layers) - 1, -1, -1):
        wname, wd, up, (hin, win), (sh, sw), (cph, cpw), (KH, KW), Ci, Co, s_bn = layers[idx]
        dc = engine.bn_bwd(ctx, s_bn, d)
        dwc = ops.new((Co, Ci, KH, KW), dc)
        ops.conv_wgrad(up, dc, dwc, KH, KW, cph, cpw, False)
        gw, acc = ctx.G(wname)
        ops_prior.tconv_weight2d_bwd(dwc, gw, acc)
        if idx == 0 and not need_dprior:
            return None
        dup = ops.conv_dgrad(dc, wd, (up.shape[1], up.shape[2]), Ci, KH, KW, cph, cpw)
        d = ops_prior.zero_upsample2d_bwd(dup, hin, win, sh, sw)
    return d if nhwc_out else ops.nhwc_to_nchw(d)                    # [B,1,26,37] or the reference's [B,37,1,26]


def _res_block_fwd(ctx: engine.Ctx, x: torch.Tensor, pre: str, text: torch.Tensor):
    B, H, W, C = x.shape
    Ct = text.shape[-1]
    c1, s_c1 = engine.conv_fwd(ctx, x, pre + 'conv1.weight', pre + 'conv1.bias', 3, 1)
    n1, s_bn1 = engine.bn_fwd(ctx, c1, pre + 'bn1.', ACT_NONE)
    a1 = ops.act_fwd(n1, ACT_PRELU, ctx.P(pre + 'prelu.weight'))
    cat = ops.new((B, H, W, C + Ct), x)
    ops.copy_cols(a1, C, cat, C + Ct, B * H * W, C)
    ops.copy_cols(text, Ct, cat, C + Ct, B * H * W, Ct, dst_off=C)
    c2, s_c2 = engine.conv_fwd(ctx, cat, pre + 'conv2.weight', pre + 'conv2.bias', 3, 1)
    out, s_bn2 = engine.bn_fwd(ctx, c2, pre + 'bn2.', ACT_NONE, residual=x)
    return out, (s_c1, s_bn1, n1, s_c2, s_bn2, pre, (B, H, W, C, Ct))


def _res_block_bwd(ctx: engine.Ctx, saved, dout: torch.Tensor, dtext: torch.Tensor) -> torch.Tensor:
    """dout is consumed (the skip gradient is accumulated into it); returns d(block input)."""
    s_c1, s_bn1, n1, s_c2, s_bn2, pre, (B, H, W, C, Ct) = saved
    M = B * H * W
    dc2 = engine.bn_bwd(ctx, s_bn2, dout)
    dcat = engine.conv_bwd(ctx, s_c2, dc2)                           # [B,H,W,C+Ct]
    ops.copy_cols(dcat, C + Ct, dtext, Ct, M, Ct, accumulate=True, src_off=C)
    da1 = ops.new((B, H, W, C), dout)
    ops.copy_cols(dcat, C + Ct, da1, C, M, C)
    gs, acc = ctx.G(pre + 'prelu.weight')
    dn1 = ops.act_bwd(da1, n1, ACT_PRELU, ctx.P(pre + 'prelu.weight'), gs, acc)
    dc1 = engine.bn_bwd(ctx, s_bn1, dn1)
    return engine.conv_bwd(ctx, s_c1, dc1, True, dx_out=dout, dx_accumulate=True)


def srresnet_fwd(ctx: engine.Ctx, x_nchw: torch.Tensor, text: torch.Tensor):
    """(sr NCHW [B,C,2H,2W], saved); ``text`` is the NHWC prior map [B,H,W,32]."""
    x = ops.nchw_to_nhwc(x_nchw)
    b1pre, s_b1 = engine.conv_fwd(ctx, x, 'block1.0.weight', 'block1.0.bias', 9, 4, need_dgrad=False)
    b1 = ops.act_fwd(b1pre, ACT_PRELU, ctx.P('block1.1.weight'))
    b, s_blocks = b1, []
    for i in range(2, 7):
        b, s = _res_block_fwd(ctx, b, f'block{i}.', text)
        s_blocks.append(s)
    c7, s_c7 = engine.conv_fwd(ctx, b, 'block7.0.weight', 'block7.0.bias', 3, 1)
    s8, s_bn7 = engine.bn_fwd(ctx, c7, 'block7.1.', ACT_NONE, residual=b1)
    u, s_up = engine.conv_fwd(ctx, s8, 'block8.0.conv.weight', 'block8.0.conv.bias', 3, 1)
    ua = ops.act_fwd(u, ACT_PRELU, ctx.P('block8.0.prelu.weight'))   # one shared slope: PReLU commutes with the pixel shuffle
    v = ops_prior.pixelshuffle_fwd(ua)
    sr, s_out = engine.conv_fwd(ctx, v, 'block8.1.weight', 'block8.1.bias', 9, 4, act=ACT_TANH, out_nchw=True)
    return sr, (s_b1, b1pre, s_blocks, s_c7, s_bn7, s_up, u, s_out, sr)


def srresnet_bwd(ctx: engine.Ctx, saved, dsr_nchw: torch.Tensor, dtext: torch.Tensor) -> None:
    s_b1, b1pre, s_blocks, s_c7, s_bn7, s_up, u, s_out, sr = saved
    dpre = ops.nchw_to_nhwc(dsr_nchw, tanh_out=sr)
    dv = engine.conv_bwd(ctx, s_out, dpre)
    dua = ops_prior.pixelshuffle_bwd(dv)
    gs, acc = ctx.G('block8.0.prelu.weight')
    du = ops.act_bwd(dua, u, ACT_PRELU, ctx.P('block8.0.prelu.weight'), gs, acc)
    ds8 = engine.conv_bwd(ctx, s_up, du)                             # = d(b1 + bn7(c7))
    dc7 = engine.bn_bwd(ctx, s_bn7, ds8)
    db = engine.conv_bwd(ctx, s_c7, dc7)
    for s in reversed(s_blocks):
        db = _res_block_bwd(ctx, s, db, dtext)
    db1 = ops.add(db, ds8)
    gs, acc = ctx.G('block1.1.weight')
    db1pre = ops.act_bwd(db1, b1pre, ACT_PRELU, ctx.P('block1.1.weight'), gs, acc)
    engine.conv_bwd(ctx, s_b1, db1pre, need_dx=False)


class _SRResNetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, text_emb, anchor, want_backward):
        sr, saved = module._run_forward(x, text_emb, want_backward)
        ctx.module, ctx.saved = module, saved
        ctx.text_requires_grad = text_emb is not None and text_emb.requires_grad
        return sr

    @staticmethod
    def backward(ctx, dsr):
        dtext = ctx.module._run_backward(ctx.saved, dsr, ctx.text_requires_grad)
        ctx.saved = None
        return None, None, dtext, None, None


class SRResNet_TL(nn.Module):
    def __init__(self, scale_factor=2, width=128, height=32, STN=False, mask=False, word_vec_d=300, text_emb=37, out_text_channels=32):
        super().__init__()
        if STN:
            raise NotImplementedError('SRResNet_TL(STN=True) uses the recogniser STN head, which the reference itself cannot run on a '
                                      '16x64 input (model/srresnet.py:8-9, model/recognizer/stn_head.py:33-44)')
        if scale_factor != 2:
            raise NotImplementedError('tpgsr_b200 kernels are specialised for scale_factor=2 (the reference default)')
        upsample_block_num = int(math.log(scale_factor, 2))
        in_planes = 4 if mask else 3
        self.in_planes = in_planes
        self.emb_cls = text_emb
        self.block1 = nn.Sequential(nn.Conv2d(in_planes, 64, kernel_size=9, padding=4), nn.PReLU())
        for i in range(2, 7):
            setattr(self, f'block{i}', ResidualBlock_TL(64, out_text_channels))
        self.block7 = nn.Sequential(nn.Conv2d(64, 64, kernel_size=3, padding=1), nn.BatchNorm2d(64))
        block8 = [UpsampleBLock(64, 2) for _ in range(upsample_block_num)]
        block8.append(nn.Conv2d(64, in_planes, kernel_size=9, padding=4))
        self.block8 = nn.Sequential(*block8)
        self.tps_inputsize = [height // scale_factor, width // scale_factor]
        self.stn = False
        self.infoGen = InfoGen(text_emb, out_text_channels)
        self._flat = FlatState(self, ())

    def _run_forward(self, x, text_emb, want_backward: bool):
        if not x.is_cuda:
            raise RuntimeError('tpgsr_b200 modules run on CUDA (sm_100a) only — there is no CPU fallback')
        if x.dtype != torch.float32 or x.dim() != 4 or x.shape[1] != self.in_planes:
            raise ValueError(f'expected float32 NCHW input with {self.in_planes} channels, got {tuple(x.shape)} {x.dtype}')
        st = self._flat
        st.ensure()
        train_bwd = want_backward and any(p.requires_grad for p in self.parameters())
        ectx = engine.Ctx(st.views, st.gviews if train_bwd else None, self.training)
        if self.training and st.nbt is not None:
            ops.add_i64(st.nbt, 1)
        B, _, H, W = x.shape
        if text_emb is None:
            text_emb = torch.zeros((B, self.emb_cls, 1, 26), dtype=torch.float32, device=x.device)
        if not text_emb.is_cuda:
            raise RuntimeError('tpgsr_b200 runs on CUDA only — text_emb must live on the same device as the images')
        text, s_text = info_gen2d_fwd(ectx, text_emb.contiguous(), (H, W))
        sr, saved = srresnet_fwd(ectx, x, text)
        return sr, (ectx, saved, s_text) if train_bwd else None

    def _run_backward(self, saved_all, dsr, text_requires_grad: bool):
        if saved_all is None:
            raise RuntimeError('backward through an SRResNet_TL forward that ran without grad enabled')
        ectx, saved, s_text = saved_all
        st = self._flat
        ectx._written = set(st.live_grads())
        if not dsr.is_contiguous():
            dsr = dsr.contiguous()
        dtext = torch.zeros_like(s_text['text'])
        srresnet_bwd(ectx, saved, dsr, dtext)
        dprior = info_gen2d_bwd(ectx, s_text, dtext, text_requires_grad)
        ectx.join()
        params = dict(self.named_parameters())
        for name in ectx._written:
            p = params.get(name)
            if p is not None and p.requires_grad:
                p.grad = st.gviews[name]
        return dprior

    @property
    def module(self):
        """The un-wrapped model (``save_checkpoint`` reads ``netG.module``, interfaces/base.py:556; one process per GPU has no wrapper)."""
        return self

    def complexity_info(self, input_res):
        """Multiply-accumulates of the eval-mode forward for one ``(C, H, W)`` image (what ``ptflops`` prints, interfaces/base.py:376)."""
        c_in, H, W = input_res
        hw, c, t_ch = H * W, 64, 32
        macs = c_in * c * 81 * hw + 5 * (c * c * 9 + (c + t_ch) * c * 9) * hw + c * c * 9 * hw + c * 4 * c * 9 * hw + c * c_in * 81 * 4 * hw
        h, w = 1, 26
        for i, (s_, p_) in enumerate([((2, 2), (0, 0)), ((2, 2), (0, 0)), ((2, 2), (1, 1)), ((2, 1), (1, 1))], 1):
            wt = getattr(self.infoGen, f'tconv{i}').weight
            macs += h * w * wt.shape[0] * wt.shape[1] * 9
            h, w = (h - 1) * s_[0] - 2 * p_[0] + 3, (w - 1) * s_[1] - 2 * p_[1] + 3
        return macs

    def forward(self, x, text_emb=None):
        anchor = next(self.parameters())
        return _SRResNetFunction.apply(self, x, text_emb, anchor, torch.is_grad_enabled())
