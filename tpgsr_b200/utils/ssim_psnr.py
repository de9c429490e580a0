"""PSNR / SSIM drop-ins (reference: ``utils/ssim_psnr.py:9-15`` ``calculate_psnr``, ``:53-78`` ``SSIM``) on one fused CUDA kernel.

Same call surface as the reference (``interfaces/base.py:230-231`` binds ``cal_psnr = ssim_psnr.calculate_psnr`` and
``cal_ssim = ssim_psnr.SSIM()``): both take two NCHW batches in [0, 1], compare the first three channels and return a 0-dim
device tensor.  One launch computes both metrics; the second call on the same pair is served from a one-entry cache."""
from __future__ import annotations

import weakref

import torch
from torch import nn

from .. import ops
from ..ops import _call, ptr, query, stream, workspace

_last = {}


def _both(img1: torch.Tensor, img2: torch.Tensor) -> torch.Tensor:
    if not (img1.is_cuda and img2.is_cuda):
        raise RuntimeError('tpgsr_b200 metrics run on CUDA only — there is no CPU fallback')
    if img1.shape != img2.shape or img1.dim() != 4 or img1.shape[1] < 3:
        raise ValueError(f'expected two NCHW batches with >= 3 channels, got {tuple(img1.shape)} and {tuple(img2.shape)}')
    # one-entry cache for the reference's call pattern cal_psnr(sr, hr); cal_ssim(sr, hr) on the SAME tensor objects
    # (interfaces/super_resolution.py:1011-1012).  Identity + version, not addresses: the next batch's tensors usually land on the same
    # addresses with the same version counter, and must not be served the previous batch's result.
    hit = _last.get('hit')
    if hit is not None and hit[0]() is img1 and hit[1]() is img2 and hit[2] == (img1._version, img2._version):
        return hit[3]
    a, b = img1.detach().float(), img2.detach().float()
    B, _, H, W = a.shape
    out = torch.empty(2, dtype=torch.float32, device=a.device)
    wp, wn = workspace(query('tpgsr_psnr_ssim_workspace_bytes', B, 3), a.device)
    _call('tpgsr_psnr_ssim', ptr(a), ptr(b), B, 3, H, W, *a.stride(), *b.stride(), ptr(out), wp, wn, stream())
    _last['hit'] = (weakref.ref(img1), weakref.ref(img2), (img1._version, img2._version), out)
    return out


def calculate_psnr(img1: torch.Tensor, img2: torch.Tensor) -> torch.Tensor:
    """20 * log10(255 / sqrt(mse)) over the RGB channels, images in [0, 1] (inf for identical images)."""
    return _both(img1, img2)[0]


class SSIM(nn.Module):
    def __init__(self, window_size: int = 11, size_average: bool = True):
        super().__init__()
        if window_size != 11 or not size_average:
            raise NotImplementedError('tpgsr_b200 SSIM implements the reference default: window 11, sigma 1.5, global mean')
        self.window_size, self.size_average = window_size, size_average

    def forward(self, img1: torch.Tensor, img2: torch.Tensor) -> torch.Tensor:
        return _both(img1, img2)[1]
