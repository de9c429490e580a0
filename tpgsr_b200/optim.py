"""Fused flat-buffer optimiser: global-norm clip + Adam in two kernels (reference: ``clip_grad_norm_`` +
``optim.Adam`` at interfaces/super_resolution.py:422-424, interfaces/base.py:449-450)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import ops
from .ops import _call, ptr, query, stream, workspace


class FlatAdam:
    """Adam(lr, betas, eps) over one flat fp32 buffer; gradients in ``[0, n_clipped)`` are first scaled by
    ``min(1, max_norm / (||g[0:n_clipped]|| + 1e-6))`` (torch's clip_grad_norm_ semantics).  Step counter and bias
    corrections live on the device (``state``), so ``step()`` is CUDA-graph replayable.

    state = [step, 1-beta1^t, 1-beta2^t, last grad norm, clip coefficient, ...]."""

    def __init__(self, params: torch.Tensor, n_clipped: Optional[int] = None, lr: float = 1e-3,
                 betas: Tuple[float, float] = (0.5, 0.999), eps: float = 1e-8, max_norm: float = 0.25,
                 grads: Optional[torch.Tensor] = None):
        assert params.is_cuda and params.dtype == torch.float32 and params.is_contiguous()
        self.params = params
        self.grads = grads if grads is not None else torch.zeros_like(params)
        self.exp_avg = torch.zeros_like(params)
        self.exp_avg_sq = torch.zeros_like(params)
        self.n = params.numel()
        self.n_clipped = self.n if n_clipped is None else int(n_clipped)
        self.lr, self.betas, self.eps, self.max_norm = lr, betas, eps, max_norm
        self.state = torch.zeros(8, dtype=torch.float32, device=params.device)

    def step(self, grad_scale: float = 1.0) -> None:
        wp, wn = workspace(query('tpgsr_optim_workspace_bytes', max(self.n_clipped, 1)), self.params.device)
        _call('tpgsr_optim_prepare', ptr(self.grads), self.n_clipped, self.max_norm if self.n_clipped > 0 else 0.0,
              self.betas[0], self.betas[1], grad_scale, ptr(self.state), 1, wp, wn, stream())
        _call('tpgsr_adam_step', ptr(self.params), ptr(self.grads), ptr(self.exp_avg), ptr(self.exp_avg_sq), self.n,
              self.n_clipped, ptr(self.state), self.lr, self.betas[0], self.betas[1], self.eps, grad_scale, stream())

    @property
    def grad_norm(self) -> float:
        return float(self.state[3])


def clip_grad_norm_flat(grads: torch.Tensor, max_norm: float, state: Optional[torch.Tensor] = None) -> torch.Tensor:
    """In-place clip of a flat gradient buffer; returns the 8-float state tensor (norm in [3], coefficient in [4])."""
    if state is None:
        state = torch.zeros(8, dtype=torch.float32, device=grads.device)
    wp, wn = workspace(query('tpgsr_optim_workspace_bytes', grads.numel()), grads.device)
    _call('tpgsr_optim_prepare', ptr(grads), grads.numel(), max_norm, 0.5, 0.999, 1.0, ptr(state), 0, wp, wn, stream())
    _call('tpgsr_scale_inplace', ptr(grads), grads.numel(), state.data_ptr() + 16, 1.0, stream())
    return state
