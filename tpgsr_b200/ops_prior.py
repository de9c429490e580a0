"""Tensor-level wrappers for the text-prior branch kernels (resampling, prior formatting, LSTM recurrence)."""
from __future__ import annotations

import os
from typing import Optional

import torch

from ._lib import LIB
from .ops import _call, new, ptr, stream


# ------------------------------------------------------------------------------------------------ parse_crnn_data
def bicubic_gray_fwd(imgs: torch.Tensor, out_hw=(32, 100)) -> torch.Tensor:
    """interfaces/base.py:806-829.  imgs: NCHW (any strides, >=3 channels) -> gray [B,1,32,100]."""
    B, C, Hi, Wi = imgs.shape
    assert C >= 3
    out = new((B, 1, out_hw[0], out_hw[1]), imgs)
    sb, sc, sh, sw = imgs.stride()
    _call('tpgsr_bicubic_gray_fwd', ptr(imgs), sb, sc, sh, sw, B, Hi, Wi, ptr(out), out_hw[0], out_hw[1], stream())
    return out


def bicubic_gray_bwd(dgray: torch.Tensor, din: torch.Tensor, accumulate: bool) -> torch.Tensor:
    """din: NCHW gradient buffer of the resized image (channels 0..2 receive the gradient)."""
    B, C, Hi, Wi = din.shape
    _, _, Ho, Wo = dgray.shape
    sb, sc, sh, sw = din.stride()
    _call('tpgsr_bicubic_gray_bwd', ptr(dgray), B, Hi, Wi, Ho, Wo, ptr(din), sb, sc, sh, sw, int(accumulate), stream())
    return din


# ------------------------------------------------------------------------------------------------ InfoGen helpers
def bilinear_rows_fwd(x: torch.Tensor, Ho: int, Wo: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x NHWC [B,1,Wi,C] -> [B,Ho,Wo,C] (align_corners=True; all rows equal)."""
    B, _, Wi, C = x.shape
    if out is None:
        out = new((B, Ho, Wo, C), x)
    _call('tpgsr_bilinear_rows_fwd', ptr(x), ptr(out), B, Wi, C, Ho, Wo, out.shape[-1], stream())
    return out


def bilinear_rows_bwd(dout: torch.Tensor, Wi: int) -> torch.Tensor:
    B, Ho, Wo, C = dout.shape
    din = new((B, 1, Wi, C), dout)
    _call('tpgsr_bilinear_rows_bwd', ptr(dout), ptr(din), B, Wi, C, Ho, Wo, C, stream())
    return din


def zero_upsample_w(x: torch.Tensor, stride: int) -> torch.Tensor:
    B, _, W, C = x.shape
    y = new((B, 1, (W - 1) * stride + 1, C), x)
    _call('tpgsr_zero_upsample_w', ptr(x), ptr(y), B, W, C, stride, stream())
    return y


def zero_upsample_w_bwd(dy: torch.Tensor, W: int, stride: int) -> torch.Tensor:
    B, _, _, C = dy.shape
    dx = new((B, 1, W, C), dy)
    _call('tpgsr_zero_upsample_w_bwd', ptr(dy), ptr(dx), B, W, C, stride, stream())
    return dx


def tconv_weight_fwd(wt: torch.Tensor) -> torch.Tensor:
    """ConvTranspose2d weight [Ci,Co,3,3] -> equivalent stride-1 conv weight OIHW [Co,Ci,1,3] (middle row, flipped)."""
    Ci, Co, KH, KW = wt.shape
    wc = new((Co, Ci, 1, KW), wt)
    _call('tpgsr_tconv_weight_fwd', ptr(wt), ptr(wc), Ci, Co, KH, KW, KH // 2, stream())
    return wc


def tconv_weight_bwd(dwc: torch.Tensor, dwt: torch.Tensor, accumulate: bool) -> None:
    Ci, Co, KH, KW = dwt.shape
    _call('tpgsr_tconv_weight_bwd', ptr(dwc), ptr(dwt), Ci, Co, KH, KW, KH // 2, int(accumulate), stream())


# ------------------------------------------------------------------------------------------------ prior formatting
def softmax_prior_fwd(logits: torch.Tensor, drop_first: int, need_prior: bool = True):
    """logits [T,B,C] -> (probs [T,B,C], prior NHWC [B,1,T,C] with the first `drop_first` samples zeroed)."""
    T, B, C = logits.shape
    probs = torch.empty_like(logits)
    prior = new((B, 1, T, C), logits) if need_prior else None
    _call('tpgsr_softmax_prior_fwd', ptr(logits), ptr(probs), ptr(prior), T, B, C, drop_first, stream())
    return probs, prior


def softmax_prior_bwd(probs: torch.Tensor, dprobs: Optional[torch.Tensor], dprior: Optional[torch.Tensor], drop_first: int) -> torch.Tensor:
    T, B, C = probs.shape
    dlogits = torch.empty_like(probs)
    _call('tpgsr_softmax_prior_bwd', ptr(probs), ptr(dprobs), ptr(dprior), ptr(dlogits), T, B, C, drop_first, stream())
    return dlogits


# ------------------------------------------------------------------------------------------------ LSTM
# developer A/B switch: TPGSR_LSTM_MODE=0 forces the FFMA cluster kernels of the LSTM recurrence (default 1: tensor-core kernels)
if os.environ.get('TPGSR_LSTM_MODE', '1') == '0':
    LIB.tpgsr_lstm_set_mode(0)


def lstm_transpose_whh(whh: torch.Tensor) -> torch.Tensor:
    """[2,1024,256] -> [2,256,1024]"""
    out = new((2, 256, 1024), whh)
    _call('tpgsr_lstm_transpose_whh', ptr(whh), ptr(out), stream())
    return out


def lstm_fwd(gx: torch.Tensor, whhT: torch.Tensor, bhh: torch.Tensor, T: int, B: int, need_stash: bool = True,
             batch_first: bool = False):
    """gx [T*B, 2048] -> y [T*B, 512], stash [T*B,2,5,256]; rows are (t,b) time-first or (b,t) batch-first."""
    y = new((T * B, 512), gx)
    stash = new((T * B, 2, 5, 256), gx) if need_stash else None
    _call('tpgsr_lstm_fwd', ptr(gx), ptr(whhT), ptr(bhh), ptr(y), ptr(stash), T, B, 256, int(batch_first), stream())
    return y, stash


def lstm_bwd(dy: torch.Tensor, stash: torch.Tensor, whh: torch.Tensor, T: int, B: int, batch_first: bool = False) -> torch.Tensor:
    dgx = new((T * B, 2048), dy)
    _call('tpgsr_lstm_bwd', ptr(dy), ptr(stash), ptr(whh), ptr(dgx), T, B, 256, int(batch_first), stream())
    return dgx


def lstm_shift_h(y: torch.Tensor, T: int, B: int, batch_first: bool = False) -> torch.Tensor:
    hprev = new((T * B, 2, 256), y)
    _call('tpgsr_lstm_shift_h', ptr(y), ptr(hprev), T, B, 256, int(batch_first), stream())
    return hprev


# ------------------------------------------------------------------------------------------------ general 2-D resampling (SRResNet_TL)
def zero_upsample2d(x: torch.Tensor, sh: int, sw: int) -> torch.Tensor:
    B, H, W, C = x.shape
    y = new((B, (H - 1) * sh + 1, (W - 1) * sw + 1, C), x)
    _call('tpgsr_zero_upsample2d', ptr(x), ptr(y), B, H, W, C, sh, sw, stream())
    return y


def zero_upsample2d_bwd(dy: torch.Tensor, H: int, W: int, sh: int, sw: int) -> torch.Tensor:
    B, _, _, C = dy.shape
    dx = new((B, H, W, C), dy)
    _call('tpgsr_zero_upsample2d_bwd', ptr(dy), ptr(dx), B, H, W, C, sh, sw, stream())
    return dx


def tconv_weight2d_fwd(wt: torch.Tensor) -> torch.Tensor:
    """ConvTranspose2d weight [Ci,Co,KH,KW] -> equivalent stride-1 conv weight OIHW [Co,Ci,KH,KW] (flipped)."""
    Ci, Co, KH, KW = wt.shape
    wc = new((Co, Ci, KH, KW), wt)
    _call('tpgsr_tconv_weight2d_fwd', ptr(wt), ptr(wc), Ci, Co, KH, KW, stream())
    return wc


def tconv_weight2d_bwd(dwc: torch.Tensor, dwt: torch.Tensor, accumulate: bool) -> None:
    Ci, Co, KH, KW = dwt.shape
    _call('tpgsr_tconv_weight2d_bwd', ptr(dwc), ptr(dwt), Ci, Co, KH, KW, int(accumulate), stream())


def bilinear2d_fwd(x: torch.Tensor, Ho: int, Wo: int) -> torch.Tensor:
    """NHWC [B,Hi,Wi,C] -> [B,Ho,Wo,C], F.interpolate(mode='bilinear', align_corners=True)."""
    B, Hi, Wi, C = x.shape
    out = new((B, Ho, Wo, C), x)
    _call('tpgsr_bilinear2d_fwd', ptr(x), ptr(out), B, Hi, Wi, C, Ho, Wo, C, stream())
    return out


def bilinear2d_bwd(dout: torch.Tensor, Hi: int, Wi: int) -> torch.Tensor:
    B, Ho, Wo, C = dout.shape
    din = new((B, Hi, Wi, C), dout)
    _call('tpgsr_bilinear2d_bwd', ptr(dout), ptr(din), B, Hi, Wi, C, Ho, Wo, C, stream())
    return din


def pixelshuffle_fwd(x: torch.Tensor) -> torch.Tensor:
    B, H, W, C4 = x.shape
    y = new((B, 2 * H, 2 * W, C4 // 4), x)
    _call('tpgsr_pixelshuffle_fwd', ptr(x), ptr(y), B, H, W, C4 // 4, stream())
    return y


def pixelshuffle_bwd(dy: torch.Tensor) -> torch.Tensor:
    B, H2, W2, C = dy.shape
    dx = new((B, H2 // 2, W2 // 2, 4 * C), dy)
    _call('tpgsr_pixelshuffle_bwd', ptr(dy), ptr(dx), B, H2 // 2, W2 // 2, C, stream())
    return dx
