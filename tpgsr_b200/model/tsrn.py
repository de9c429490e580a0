"""TSRN / TSRN_TL drop-in modules (reference: ``model/tsrn.py:18-215``) executed by hand-written sm_100a kernels.

Constructor kwargs, ``forward`` signatures, submodule names and ``state_dict`` keys are the reference's
(``interfaces/base.py:296-297,324-326`` construct them); the parameters live in reference-shaped container submodules
(never called), and ``forward`` runs ``tpgsr_b200.engine.tsrn_fwd`` / ``tsrn_bwd`` through the C ABI.  There is no
PyTorch fallback: calling a module whose parameters are not on a CUDA device raises.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
from torch import nn

from .. import engine, ops
from ._flat import FlatState
from .stn_head import STNHead
from .tps_spatial_transformer import TPSSpatialTransformer


class mish(nn.Module):
    """Parameter-free marker (model/tsrn.py:480-488); the activation itself is fused into the BN-apply kernels."""

    def __init__(self):
        super().__init__()
        self.activated = True


class GruBlock(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        assert out_channels % 2 == 0
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=1, padding=0)
        self.gru = nn.GRU(out_channels, out_channels // 2, bidirectional=True, batch_first=True)


class RecurrentResidualBlock(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv1 = nn.Conv2d(channels, channels, kernel_size=3, padding=1)
        self.bn1 = nn.BatchNorm2d(channels)
        self.gru1 = GruBlock(channels, channels)
        self.prelu = mish()
        self.conv2 = nn.Conv2d(channels, channels, kernel_size=3, padding=1)
        self.bn2 = nn.BatchNorm2d(channels)
        self.gru2 = GruBlock(channels, channels)


class RecurrentResidualBlockTL(nn.Module):
    def __init__(self, channels, text_channels):
        super().__init__()
        self.conv1 = nn.Conv2d(channels, channels, kernel_size=3, padding=1)
        self.bn1 = nn.BatchNorm2d(channels)
        self.gru1 = GruBlock(channels + text_channels, channels)
        self.prelu = mish()
        self.conv2 = nn.Conv2d(channels, channels, kernel_size=3, padding=1)
        self.bn2 = nn.BatchNorm2d(channels)
        self.gru2 = GruBlock(channels, channels)


class UpsampleBLock(nn.Module):
    def __init__(self, in_channels, up_scale):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, in_channels * up_scale ** 2, kernel_size=3, padding=1)
        self.pixel_shuffle = nn.PixelShuffle(up_scale)
        self.prelu = mish()


class InfoGen(nn.Module):
    def __init__(self, t_emb, output_size):
        super().__init__()
        self.tconv1 = nn.ConvTranspose2d(t_emb, 512, 3, 2, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(512)
        self.tconv2 = nn.ConvTranspose2d(512, 128, 3, 2, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(128)
        self.tconv3 = nn.ConvTranspose2d(128, 64, 3, 2, padding=1, bias=False)
        self.bn3 = nn.BatchNorm2d(64)
        self.tconv4 = nn.ConvTranspose2d(64, output_size, 3, (2, 1), padding=(1, 0), bias=False)
        self.bn4 = nn.BatchNorm2d(output_size)


def _gru_fusions(prefix: str, hidden: int, in_ch: int):
    g = prefix + 'gru.'
    return [
        (g + 'weight_ih', [g + 'weight_ih_l0', g + 'weight_ih_l0_reverse'], (6 * hidden, in_ch)),
        (g + 'weight_hh', [g + 'weight_hh_l0', g + 'weight_hh_l0_reverse'], (2, 3 * hidden, hidden)),
        (g + 'bias_ih', [g + 'bias_ih_l0', g + 'bias_ih_l0_reverse'], (6 * hidden,)),
        (g + 'bias_hh', [g + 'bias_hh_l0', g + 'bias_hh_l0_reverse'], (2, 3 * hidden)),
    ]


class _TSRNFunction(torch.autograd.Function):
    """Autograd hand-off: one node for the whole network.  Parameter gradients are written straight into the module's
    flat gradient buffer (and published as ``.grad``); only d/d(text prior) is returned to autograd."""

    @staticmethod
    def forward(ctx, module, x, text_emb, anchor, want_backward):
        sr, saved = module._run_forward(x, text_emb, want_backward=want_backward)
        ctx.module, ctx.saved = module, saved
        ctx.text_requires_grad = text_emb is not None and text_emb.requires_grad
        return sr

    @staticmethod
    def backward(ctx, dsr):
        dtext = ctx.module._run_backward(ctx.saved, dsr, ctx.text_requires_grad)
        ctx.saved = None
        return None, None, dtext, None, None


class _TSRNBase(nn.Module):
    _tl = False

    def _setup(self, scale_factor, width, height, STN, srb_nums, mask, hidden_units, text_channels=0):
        in_planes = 4 if mask else 3
        assert math.log(scale_factor, 2) % 1 == 0
        if scale_factor != 2 or hidden_units != 32:
            raise NotImplementedError('tpgsr_b200 kernels are specialised for scale_factor=2, hidden_units=32 (the reference defaults)')
        c = 2 * hidden_units
        self.block1 = nn.Sequential(nn.Conv2d(in_planes, c, kernel_size=9, padding=4), nn.PReLU())
        self.srb_nums = srb_nums
        for i in range(srb_nums):
            blk = RecurrentResidualBlockTL(c, text_channels) if self._tl else RecurrentResidualBlock(c)
            setattr(self, 'block%d' % (i + 2), blk)
        return in_planes, c

    def _finish(self, in_planes, c, scale_factor, width, height, STN, srb_nums):
        setattr(self, 'block%d' % (srb_nums + 2), nn.Sequential(nn.Conv2d(c, c, kernel_size=3, padding=1), nn.BatchNorm2d(c)))
        tail = [UpsampleBLock(c, 2) for _ in range(int(math.log(scale_factor, 2)))]
        tail.append(nn.Conv2d(c, in_planes, kernel_size=9, padding=4))
        setattr(self, 'block%d' % (srb_nums + 3), nn.Sequential(*tail))
        self.tps_inputsize = [height // scale_factor, width // scale_factor]
        self.stn = STN
        self.in_planes = in_planes
        if self.stn:
            self.tps = TPSSpatialTransformer(output_image_size=tuple(self.tps_inputsize), num_control_points=20,
                                             margins=(0.05, 0.05))
            self.stn_head = STNHead(in_planes=in_planes, num_ctrlpoints=20, activation='none', input_size=self.tps_inputsize)
        fused = []
        for i in range(srb_nums):
            fused += _gru_fusions(f'block{i + 2}.gru1.', 32, c)
            fused += _gru_fusions(f'block{i + 2}.gru2.', 32, c)
        self._flat = FlatState(self, fused)

    # ---------------------------------------------------------------------------------------------------- execution
    def _check_input(self, x):
        if not x.is_cuda:
            raise RuntimeError('tpgsr_b200 modules run on CUDA (sm_100a) only — there is no CPU fallback')
        if x.dtype != torch.float32 or x.dim() != 4 or x.shape[1] != self.in_planes:
            raise ValueError(f'expected float32 NCHW input with {self.in_planes} channels, got {tuple(x.shape)} {x.dtype}')
        if x.requires_grad and torch.is_grad_enabled():
            raise NotImplementedError('TSRN input-image gradients are not produced (the reference feeds data tensors; the cascade without '
                                      '--sr_share is out of scope) — detach the image or use tpgsr_b200.train.TPGSRTrainStep')

    def _run_forward(self, x, text_emb, want_backward: bool):
        self._check_input(x)
        st = self._flat
        st.ensure()
        train_bwd = want_backward and any(p.requires_grad for p in self.parameters())
        ectx = engine.Ctx(st.views, st.gviews if train_bwd else None, self.training)
        if self.training and st.nbt is not None:
            self._bump_bn_counters(st)
        text, s_text = (None, None)
        if self._tl:
            text, s_text = self._expand_prior(ectx, x, text_emb)
        sr, saved = engine.tsrn_fwd(ectx, x, self.stn, self.srb_nums, text)
        return sr, (ectx, saved, s_text) if train_bwd else None

    def _bump_bn_counters(self, st):
        # every BN of this module sees exactly one batch per training forward, except the STN head's in eval mode
        ops.add_i64(st.nbt, 1)

    def _run_backward(self, saved_all, dsr, text_requires_grad: bool):
        if saved_all is None:
            raise RuntimeError('no saved state for this backward: either the forward ran without grad enabled, or this is a second backward '
                               'through the same forward (retain_graph / double backward are not supported: the stash is released after use)')
        ectx, saved, s_text = saved_all
        st = self._flat
        ectx._written = set(st.live_grads())
        if not dsr.is_contiguous():
            dsr = dsr.contiguous()
        dtext_map = None
        if self._tl:
            dtext_map = torch.zeros_like(s_text['text'])
        engine.tsrn_bwd(ectx, saved, dsr, dtext_map)
        dprior = None
        if self._tl:
            dprior = self._expand_prior_bwd(ectx, s_text, dtext_map, text_requires_grad)
        ectx.join()
        self._publish(ectx)
        return dprior

    def _publish(self, ectx):
        st = self._flat
        members = {f: names for f, names, _ in st.fused}
        params = dict(self.named_parameters())
        for name in ectx._written:
            for n in members.get(name, [name]):
                p = params.get(n)
                if p is not None and p.requires_grad:
                    p.grad = st.gviews[n]

    # ---------------------------------------------------------------------------------------------------- caller conveniences
    @property
    def module(self):
        """``save_checkpoint`` reads ``netG.module.state_dict()`` when the model was wrapped in ``nn.DataParallel``
        (interfaces/base.py:556); with one process per GPU there is no wrapper, so the model is its own ``.module``."""
        return self

    def complexity_info(self, input_res):
        """Multiply-accumulates of the eval-mode forward for one ``input_res = (C, H, W)`` image — what ``ptflops`` prints for the
        reference module (interfaces/base.py:376); analytic, because the fused CUDA path never runs the ``nn.Conv2d`` children."""
        c_in, H, W = input_res
        hw, c = H * W, 64
        gru = 2 * 96 * (c + 32)                                # per position: 2 directions x 3 gates x 32 units x (64 inputs + 32 hidden)
        macs = c_in * c * 81 * hw                              # block1 9x9
        t_ch = self.infoGen.tconv4.weight.shape[1] if self._tl else 0
        for _ in range(self.srb_nums):
            macs += 2 * c * c * 9 * hw                         # conv1, conv2
            macs += ((c + t_ch) * c + gru) * hw                # gru1: 1x1 conv on [features ; prior] + BiGRU
            macs += (c * c + gru) * hw                         # gru2
        macs += c * c * 9 * hw + c * 4 * c * 9 * hw            # block7 conv, upsample conv
        macs += c * c_in * 81 * 4 * hw                         # final 9x9 on the 2x map
        if self._tl:
            w_in = 26
            for i in range(1, 5):
                wt = getattr(self.infoGen, f'tconv{i}').weight
                macs += w_in * wt.shape[0] * wt.shape[1] * 9   # transposed conv: every input position feeds 9 taps
                w_in = (w_in - 1) * (2 if i < 4 else 1) - (2 if i < 4 else 0) + 3
        return macs

    def _forward(self, x, text_emb=None):
        anchor = next(self.parameters())
        return _TSRNFunction.apply(self, x, text_emb, anchor, torch.is_grad_enabled())


class TSRN(_TSRNBase):
    """model/tsrn.py:18-78.  ``forward(x)``: x [B,4,16,64] fp32 NCHW on CUDA -> tanh output [B,4,32,128]."""

    def __init__(self, scale_factor=2, width=128, height=32, STN=False, srb_nums=5, mask=True, hidden_units=32):
        super().__init__()
        in_planes, c = self._setup(scale_factor, width, height, STN, srb_nums, mask, hidden_units)
        self._finish(in_planes, c, scale_factor, width, height, STN, srb_nums)

    def forward(self, x):
        return self._forward(x)


class TSRN_TL(_TSRNBase):
    """model/tsrn.py:111-215.  ``forward(x, text_emb=None)``; text_emb [B,37,1,26] (softmaxed recogniser output)."""
    _tl = True

    def __init__(self, scale_factor=2, width=128, height=32, STN=False, srb_nums=5, mask=True, hidden_units=32,
                 word_vec_d=300, text_emb=37, out_text_channels=32):
        super().__init__()
        in_planes, c = self._setup(scale_factor, width, height, STN, srb_nums, mask, hidden_units, out_text_channels)
        self.feature_enhancer = None
        self.infoGen = InfoGen(text_emb, out_text_channels)
        self.emb_cls = text_emb
        self._finish(in_planes, c, scale_factor, width, height, STN, srb_nums)

    def forward(self, x, text_emb=None):
        return self._forward(x, text_emb)

    def _expand_prior(self, ectx, x, text_emb):
        from ..infogen import info_gen_fwd
        return info_gen_fwd(ectx, x, text_emb, self.emb_cls)

    def _expand_prior_bwd(self, ectx, s_text, dtext_map, need_dprior):
        from ..infogen import info_gen_bwd
        return info_gen_bwd(ectx, s_text, dtext_map, need_dprior)
