"""ImageLoss / GradientPriorLoss drop-ins (reference: ``loss/image_loss.py:10-51``) on one fused CUDA kernel pair."""
from __future__ import annotations

import torch
from torch import nn

from .. import ops


class _ImageLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, out, target, gradient, w0, w1):
        if not out.is_cuda:
            raise RuntimeError('tpgsr_b200 losses run on CUDA only — there is no CPU fallback')
        out_c = out if out.is_contiguous() else out.contiguous()
        tgt_c = target if target.is_contiguous() else target.contiguous()
        loss = ops.image_loss_fwd(out_c, tgt_c, gradient, w0, w1)
        ctx.save_for_backward(out_c, tgt_c)
        ctx.cfg = (gradient, w0, w1)
        return loss.view(())

    @staticmethod
    def backward(ctx, dloss):
        out_c, tgt_c = ctx.saved_tensors
        gradient, w0, w1 = ctx.cfg
        dout = ops.image_loss_bwd(out_c, tgt_c, gradient, w0, w1, 1.0, upstream=dloss.contiguous().view(1))
        return dout, None, None, None, None


class GradientPriorLoss(nn.Module):
    """mean |gm(out) - gm(target)| with gm = central-difference gradient magnitude (loss/image_loss.py:33-51)."""

    def forward(self, out_images, target_images):
        # the reference applies the gradient map to EVERY channel of what it is given (ImageLoss hands it out[:, :3]); the kernel's
        # gradient term covers up to three channels per call, so wider inputs are walked in groups of three (weighted back to one mean)
        C = out_images.shape[1]
        if C <= 3:
            return _ImageLossFn.apply(out_images, target_images, True, 0.0, 1.0)
        total = 0.0
        for a in range(0, C, 3):
            b = min(a + 3, C)
            total = total + _ImageLossFn.apply(out_images[:, a:b], target_images[:, a:b], True, 0.0, 1.0) * ((b - a) / C)
        return total


class ImageLoss(nn.Module):
    def __init__(self, gradient=True, loss_weight=[20, 1e-4]):
        super().__init__()
        if gradient:
            self.GPLoss = GradientPriorLoss()
        self.gradient = gradient
        self.loss_weight = loss_weight

    def forward(self, out_images, target_images, grad_mask=None):
        # grad_mask is accepted and ignored, exactly like the reference (loss/image_loss.py:21-23)
        if self.gradient and out_images.shape[1] < 3:
            raise ValueError('gradient prior needs at least 3 channels')
        return _ImageLossFn.apply(out_images, target_images, bool(self.gradient), float(self.loss_weight[0]),
                                  float(self.loss_weight[1]) if self.gradient else 0.0)
