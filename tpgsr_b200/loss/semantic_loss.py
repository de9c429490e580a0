"""SemanticLoss drop-in (reference: ``loss/semantic_loss.py:10-38``): mean|gt-pred| + KLDiv(log(pred+1e-20), gt+1e-20)
with the legacy element-mean reduction, fused into one forward and one backward kernel."""
from __future__ import annotations

import torch
from torch import nn

from .. import ops


class _SemanticLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, gt):
        if not pred.is_cuda:
            raise RuntimeError('tpgsr_b200 losses run on CUDA only — there is no CPU fallback')
        p = pred if pred.is_contiguous() else pred.contiguous()
        g = gt if gt.is_contiguous() else gt.contiguous()
        loss = ops.semantic_loss_fwd(p, g)
        ctx.save_for_backward(p, g)
        return loss.view(())

    @staticmethod
    def backward(ctx, dloss):
        p, g = ctx.saved_tensors
        return ops.semantic_loss_bwd(p, g, 1.0, upstream=dloss.contiguous().view(1)), None


class SemanticLoss(nn.Module):
    def __init__(self, margin=0.1):
        super().__init__()
        self.margin = margin
        self.lambda1 = 1.0
        self.lambda2 = 1.0

    def forward(self, pred_vec, gt_vec):
        return _SemanticLossFn.apply(pred_vec, gt_vec)
