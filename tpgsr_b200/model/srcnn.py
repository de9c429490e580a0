"""SRCNN drop-in module (reference: ``model/srcnn.py:109-145``; BASELINE config 1) executed by the sm_100a kernels.

``SRCNN(scale_factor, width, height, STN)`` keeps the reference's constructor, ``forward(x)`` signature and ``state_dict`` keys
(``conv{1,2,3}.{weight,bias}``).  ``forward``: nearest x2 (fused with the NCHW->NHWC boundary change) -> conv9x9(3->64)+ReLU ->
conv1x1(64->32)+ReLU -> conv5x5(32->3).  The reference's SRCNN imports the *recogniser* STN head, which cannot run on a 16x64
input; like the shipped ``train_SRCNN.sh`` this module supports ``STN=False`` only.
"""
from __future__ import annotations

import torch
from torch import nn

from .. import engine, ops
from ..ops import ACT_RELU
from ._flat import FlatState


class _SRCNNFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, anchor, want_backward):
        sr, saved = module._run_forward(x, want_backward)
        ctx.module, ctx.saved = module, saved
        return sr

    @staticmethod
    def backward(ctx, dsr):
        ctx.module._run_backward(ctx.saved, dsr)
        ctx.saved = None
        return None, None, None, None


def srcnn_fwd(ectx: engine.Ctx, x_nchw: torch.Tensor, scale: int):
    """(sr NCHW [B,C,sH,sW], saved).  x_nchw may be a channel-sliced view (callers pass ``images_lr[:, :3]``)."""
    x = ops.upsample_nearest_to_nhwc(x_nchw, scale)
    a1, s1 = engine.conv_fwd(ectx, x, 'conv1.weight', 'conv1.bias', 9, 4, act=ACT_RELU, need_dgrad=False)
    a2, s2 = engine.conv_fwd(ectx, a1, 'conv2.weight', 'conv2.bias', 1, 0, act=ACT_RELU)
    sr, s3 = engine.conv_fwd(ectx, a2, 'conv3.weight', 'conv3.bias', 5, 2, out_nchw=True)
    return sr, (s1, a1, s2, a2, s3)


def srcnn_bwd(ectx: engine.Ctx, saved, dsr_nchw: torch.Tensor) -> None:
    s1, a1, s2, a2, s3 = saved
    d3 = ops.nchw_to_nhwc(dsr_nchw)
    da2 = engine.conv_bwd(ectx, s3, d3)
    d2 = ops.act_bwd(da2, a2, ACT_RELU)
    da1 = engine.conv_bwd(ectx, s2, d2)
    d1 = ops.act_bwd(da1, a1, ACT_RELU)
    engine.conv_bwd(ectx, s1, d1, need_dx=False)


class SRCNN(nn.Module):
    def __init__(self, scale_factor=2, in_planes=3, STN=False, height=32, width=128):
        super().__init__()
        if STN:
            raise NotImplementedError('SRCNN(STN=True) uses the recogniser STN head, which the reference itself cannot run on a '
                                      '16x64 input (model/srcnn.py:15-16, model/recognizer/stn_head.py:33-44)')
        self.upscale_factor = scale_factor
        self.in_planes = in_planes
        self.conv1 = nn.Conv2d(in_planes, 64, kernel_size=9, padding=4)
        self.relu1 = nn.ReLU()
        self.conv2 = nn.Conv2d(64, 32, kernel_size=1, padding=0)
        self.relu2 = nn.ReLU()
        self.conv3 = nn.Conv2d(32, in_planes, kernel_size=5, padding=2)
        self.tps_inputsize = [height // scale_factor, width // scale_factor]
        self.stn = False
        self._flat = FlatState(self, ())

    def _run_forward(self, x, want_backward: bool):
        if not x.is_cuda:
            raise RuntimeError('tpgsr_b200 modules run on CUDA (sm_100a) only — there is no CPU fallback')
        if x.dtype != torch.float32 or x.dim() != 4 or x.shape[1] != self.in_planes:
            raise ValueError(f'expected float32 NCHW input with {self.in_planes} channels, got {tuple(x.shape)} {x.dtype}')
        if x.requires_grad:
            raise NotImplementedError('SRCNN input gradients are not produced (the reference feeds data tensors)')
        st = self._flat
        st.ensure()
        train_bwd = want_backward and any(p.requires_grad for p in self.parameters())
        ectx = engine.Ctx(st.views, st.gviews if train_bwd else None, self.training)
        sr, saved = srcnn_fwd(ectx, x, self.upscale_factor)
        return sr, (ectx, saved) if train_bwd else None

    def _run_backward(self, saved_all, dsr):
        if saved_all is None:
            raise RuntimeError('backward through an SRCNN forward that ran without grad enabled')
        ectx, saved = saved_all
        st = self._flat
        ectx._written = set(st.live_grads())
        if not dsr.is_contiguous():
            dsr = dsr.contiguous()
        srcnn_bwd(ectx, saved, dsr)
        ectx.join()
        params = dict(self.named_parameters())
        for name in ectx._written:
            p = params.get(name)
            if p is not None and p.requires_grad:
                p.grad = st.gviews[name]

    @property
    def module(self):
        """The un-wrapped model (``save_checkpoint`` reads ``netG.module``, interfaces/base.py:556; one process per GPU has no wrapper)."""
        return self

    def complexity_info(self, input_res):
        """Multiply-accumulates of the forward for one ``(C, H, W)`` image (what ``ptflops`` prints, interfaces/base.py:376)."""
        c_in, H, W = input_res
        hw = H * W * self.upscale_factor ** 2
        return (c_in * 64 * 81 + 64 * 32 + 32 * c_in * 25) * hw

    def forward(self, x):
        anchor = next(self.parameters())
        return _SRCNNFunction.apply(self, x, anchor, torch.is_grad_enabled())
