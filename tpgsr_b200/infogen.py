"""InfoGen: text prior [B,37,1,26] -> spatial prior map [B,16,64,32] (reference: ``model/tsrn.py:81-108,190-196``).

The four ConvTranspose2d layers act on H=1 maps, so only the middle kernel row ever lands inside the (H_out = 1) output
(SURVEY Appendix A): each layer is run as a zero-upsample along W + a stride-1 1x3 convolution with the flipped middle
row, through the same gather-GEMM kernels as every other convolution; BN+ReLU are the fused BN kernels; the final
``F.interpolate(bilinear, align_corners=True)`` from one row to 16 is a row-broadcast lerp.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import engine, ops, ops_prior
from .ops import ACT_RELU

# (stride_w, pad_w of the ConvTranspose2d) per layer, TSRN flavour (model/tsrn.py:89-101)
_LAYERS = [(2, 1), (2, 1), (2, 1), (1, 0)]


def info_gen_fwd(ctx: engine.Ctx, x_nchw: torch.Tensor, text_emb: Optional[torch.Tensor], emb_cls: int,
                 prior_nhwc: Optional[torch.Tensor] = None):
    """``text_emb`` is the reference's NCHW prior [B,37,1,26]; the fused train step passes the NHWC prior [B,1,26,37] that the
    softmax/prior kernel already produced (``prior_nhwc``) and skips the layout change."""
    B, _, H, W = x_nchw.shape
    if prior_nhwc is not None:
        text_emb = prior_nhwc
    elif text_emb is None:
        text_emb = torch.zeros((B, emb_cls, 1, 26), dtype=torch.float32, device=x_nchw.device)
    if not text_emb.is_cuda:
        raise RuntimeError('tpgsr_b200 runs on CUDA only — text_emb must live on the same device as the images')
    t = prior_nhwc if prior_nhwc is not None else ops.nchw_to_nhwc(text_emb)      # [B,1,26,37]
    saved = []
    for i, (stride, pad_t) in enumerate(_LAYERS, 1):
        wname = f'infoGen.tconv{i}.weight'
        wt = ctx.P(wname)
        Ci, Co, KH, KW = wt.shape
        wc = ops_prior.tconv_weight_fwd(wt)                         # OIHW [Co,Ci,1,3]
        wd = None
        up = ops_prior.zero_upsample_w(t, stride) if stride > 1 else t
        pad = KW - 1 - pad_t
        s_conv = None
        if Ci % 64 == 0 and 2 * pad == KW - 1 and engine._flat_ok(up, Ci, Co, 1, KW, 0, pad, 0):
            # tconv2 / tconv3 (512 -> 128 -> 64 channels): the equivalent stride-1 conv runs on the tensor-core path.  Its weight is DERIVED
            # from the parameter, so it enters the engine under a synthetic name; its gradient is scattered back by tconv_weight_bwd.
            dname = wname + '#conv'
            ctx.params[dname] = wc
            ctx._packed_bf16.pop(dname, None)
            c, s_conv = engine.conv_fwd(ctx, up, dname, None, (1, KW), (0, pad))
        else:
            wf, wd = ops.pack_conv_weights(wc, ctx.grads is not None)
            c = ops.conv_fwd(up, wf, None, Co, 1, KW, 0, pad)
        a, s_bn = engine.bn_fwd(ctx, c, f'infoGen.bn{i}.', ACT_RELU)
        saved.append((wname, wd, up, t.shape[2], stride, pad, KW, Ci, Co, s_bn, s_conv, wc))
        t = a
    text = ops_prior.bilinear_rows_fwd(t, H, W)
    return text, {'layers': saved, 'text': text, 'w_last': t.shape[2], 'prior_shape': tuple(text_emb.shape)}


def info_gen_bwd(ctx: engine.Ctx, s_text, dtext: torch.Tensor, need_dprior: bool, nhwc_out: bool = False):
    d = ops_prior.bilinear_rows_bwd(dtext, s_text['w_last'])        # [B,1,203,32]
    layers = s_text['layers']
    for idx in range(len(layers) - 1, -1, -1):
        wname, wd, up, w_in, stride, pad, KW, Ci, Co, s_bn, s_conv, wc = layers[idx]
        dc = engine.bn_bwd(ctx, s_bn, d)
        dwc = ops.new((Co, Ci, 1, KW), dc)
        gw, acc = ctx.G(wname)
        if s_conv is not None:
            dname = wname + '#conv'
            ctx.params[dname] = wc                                  # (a shared Ctx of a later cascade stage may have re-derived it)
            ctx.grads[dname] = dwc
            ctx._written.discard(dname)                             # dwc is a fresh buffer: overwrite, the accumulation happens below
            dup = engine.conv_bwd(ctx, s_conv, dc, need_dx=not (idx == 0 and not need_dprior))
            with ctx.on_side(dwc):                                  # behind the weight gradient on the side stream (same stream: ordered)
                ops_prior.tconv_weight_bwd(dwc, gw, acc)
            if dup is None:
                return None
        else:
            ops.conv_wgrad(up, dc, dwc, 1, KW, 0, pad, False)
            ops_prior.tconv_weight_bwd(dwc, gw, acc)
            if idx == 0 and not need_dprior:
                return None
            dup = ops.conv_dgrad(dc, wd, (1, up.shape[2]), Ci, 1, KW, 0, pad)
        d = ops_prior.zero_upsample_w_bwd(dup, w_in, stride) if stride > 1 else dup
    return d if nhwc_out else ops.nhwc_to_nchw(d)                    # [B,1,26,37] or the reference's [B,37,1,26]
