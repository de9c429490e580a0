"""CRNN text recogniser forward/backward on the C ABI (reference: ``model/crnn/crnn.py:5-90``).

VGG-style conv stack (7 convs, 3 BN, 4 max-pools) -> [B,1,26,512] -> 2 x (BiLSTM(256) + Linear) -> logits [26,B,37].
Everything runs batch-first on the NHWC pixel rows (row = b*26 + t) — the LSTM kernels address (t,b) through a flag —
and only the final logits are permuted to the reference's time-first [T,B,C].
"""
from __future__ import annotations

from typing import Optional

import torch

from . import engine, ops, ops_prior
from .ops import ACT_NONE, ACT_RELU, _call, ptr, query, stream, workspace

# (name index, out channels, kernel, pad, has_bn, pool (k, s, p) or None)
_CNN = [
    (0, 64, 3, 1, False, ((2, 2), (2, 2), (0, 0))),
    (1, 128, 3, 1, False, ((2, 2), (2, 2), (0, 0))),
    (2, 256, 3, 1, True, None),
    (3, 256, 3, 1, False, ((2, 2), (2, 1), (0, 1))),
    (4, 512, 3, 1, True, None),
    (5, 512, 3, 1, False, ((2, 2), (2, 1), (0, 1))),
    (6, 512, 2, 0, True, None),
]


def _permute_rows(x2d: torch.Tensor, outer: int, inner: int) -> torch.Tensor:
    """[outer*inner, C] with rows (o,i) -> [inner*outer, C] with rows (i,o) (one strided-gather kernel)."""
    C = x2d.shape[-1]
    y = ops.new((inner * outer, C), x2d)
    # dst[i][o][c] = src[(o*inner + i)*C + c]: NCHW-view gather with H'=inner (stride C), W'=outer (stride inner*C), C'=C (stride 1)
    _call('tpgsr_nchw_to_nhwc', ptr(x2d), ptr(y), 1, C, inner, outer, 0, 1, C, inner * C, C, None, stream())
    return y


def _wgrad_pitched(x: torch.Tensor, x_off: int, ldx: int, dy: torch.Tensor, dy_off: int, ldy: int, dw: torch.Tensor, M: int,
                   Ci: int, Co: int, accumulate: bool) -> None:
    """dw [Co,Ci] (+)= dy[:, dy_off:dy_off+Co]^T @ x[:, x_off:x_off+Ci] on pitched row matrices (1x1 'conv' wgrad)."""
    nbytes = query('tpgsr_conv2d_wgrad_workspace_bytes', M, 1, 1, Ci, 1, 1, Co, 1, 1)
    wp, wn = workspace(nbytes, x.device)
    _call('tpgsr_conv2d_wgrad', x.data_ptr() + 4 * x_off, dy.data_ptr() + 4 * dy_off, ptr(dw), M, 1, 1, Ci, 1, 1, Co, 1, 1, 0, 0,
          ldx, ldy, int(accumulate), wp, wn, stream())


def _bilstm_fwd(ctx: engine.Ctx, x2d: torch.Tensor, pre: str, T: int, B: int):
    """BidirectionalLSTM (model/crnn/crnn.py:5-26): rows of x2d are (b,t)."""
    M, I = x2d.shape
    gx, s_ih = engine.conv_fwd(ctx, x2d.view(M, 1, 1, I), pre + 'rnn.weight_ih', pre + 'rnn.bias_ih', 1, 0)
    whh = ctx.P(pre + 'rnn.weight_hh')
    whhT = ops_prior.lstm_transpose_whh(whh)
    y, stash = ops_prior.lstm_fwd(gx.view(M, 2048), whhT, ctx.P(pre + 'rnn.bias_hh'), T, B, ctx.grads is not None, batch_first=True)
    e, s_emb = engine.conv_fwd(ctx, y.view(M, 1, 1, 512), pre + 'embedding.weight', pre + 'embedding.bias', 1, 0)
    return e.view(M, -1), (s_ih, s_emb, y, stash, pre, T, B)


def _bilstm_bwd(ctx: engine.Ctx, saved, de: torch.Tensor, need_dx: bool = True):
    s_ih, s_emb, y, stash, pre, T, B = saved
    M = y.shape[0]
    dy = engine.conv_bwd(ctx, s_emb, de.view(M, 1, 1, -1)).view(M, 512)
    whh = ctx.P(pre + 'rnn.weight_hh')
    dgx = ops_prior.lstm_bwd(dy, stash, whh, T, B, batch_first=True)                     # [M, 2048]
    hprev = ops_prior.lstm_shift_h(y, T, B, batch_first=True)                            # [M, 2, 256]
    gwhh, acc = ctx.G(pre + 'rnn.weight_hh')                                             # [2,1024,256]
    for d in (0, 1):
        _wgrad_pitched(hprev, d * 256, 512, dgx, d * 1024, 2048, gwhh[d], M, 256, 1024, acc)
    gbih, accb = ctx.G(pre + 'rnn.bias_ih')
    ops.colsum(dgx, gbih.view(-1), accb)
    gbhh, _ = ctx.G(pre + 'rnn.bias_hh')
    # b_ih and b_hh always receive identical gradient contributions, so they stay equal: copy instead of a second reduction
    ops.copy_cols(gbih.view(1, -1), 2048, gbhh.view(1, -1), 2048, 1, 2048, accumulate=False)
    x_ih, wname, _, kh, kw, ph, pw, as4d, ldx = s_ih
    return engine.conv_bwd(ctx, (x_ih, wname, None, kh, kw, ph, pw, as4d, ldx), dgx.view(M, 1, 1, 2048), need_dx)


def crnn_fwd(ctx: engine.Ctx, gray_nchw: torch.Tensor, need_input_grad: bool = False):
    """gray [B,1,32,100] (NCHW; C == 1 so it is also NHWC) -> (logits [26,B,37], saved)."""
    B, C, H, W = gray_nchw.shape
    x = gray_nchw.reshape(B, H, W, 1) if gray_nchw.is_contiguous() else gray_nchw.contiguous().view(B, H, W, 1)
    saved = []
    for i, co, k, pad, has_bn, pool in _CNN:
        act = ACT_NONE if has_bn else ACT_RELU
        c, s_c = engine.conv_fwd(ctx, x, f'cnn.conv{i}.weight', f'cnn.conv{i}.bias', k, pad, act=act,
                                 need_dgrad=(i > 0) or need_input_grad)
        s_bn = None
        if has_bn:
            # conv2 / conv4's BN+ReLU output feeds the next 3x3 conv directly: emit its operand planes from the apply pass
            a, s_bn = engine.bn_fwd(ctx, c, f'cnn.batchnorm{i}.', ACT_RELU, planes_P=1 if (pool is None and i < 6) else None)
        else:
            a = c
        s_pool = None
        if pool is not None:
            pk, ps, pp = pool
            p, arg = ops.maxpool_fwd(a, pk, ps, pp, need_argmax=ctx.grads is not None)
            s_pool = (arg, tuple(a.shape), pk, ps, pp)
            x = p
        else:
            x = a
        saved.append((s_c, s_bn, s_pool, a if not has_bn else None))
    Bq, Hq, T, Cq = x.shape
    if Hq != 1:
        raise ValueError('the height of conv must be 1')          # model/crnn/crnn.py:83
    seq = x.view(B * T, Cq)
    e0, s_l0 = _bilstm_fwd(ctx, seq, 'rnn.0.', T, B)
    e1, s_l1 = _bilstm_fwd(ctx, e0, 'rnn.1.', T, B)
    logits = _permute_rows(e1, B, T).view(T, B, -1)
    return logits, (saved, s_l0, s_l1, T, B, need_input_grad)


def crnn_bwd(ctx: engine.Ctx, saved_all, dlogits: torch.Tensor) -> Optional[torch.Tensor]:
    saved, s_l0, s_l1, T, B, need_input_grad = saved_all
    nc = dlogits.shape[-1]
    d = _permute_rows(dlogits.reshape(T * B, nc), T, B)                  # rows (b,t)
    d = _bilstm_bwd(ctx, s_l1, d).view(B * T, -1)
    d = _bilstm_bwd(ctx, s_l0, d).view(B, 1, T, -1)
    for idx in range(len(saved) - 1, -1, -1):
        s_c, s_bn, s_pool, relu_out = saved[idx]
        if s_pool is not None:
            arg, in_shape, pk, ps, pp = s_pool
            d = ops.maxpool_bwd(d, arg, in_shape, pk, ps, pp)
        if s_bn is not None:
            d = engine.bn_bwd(ctx, s_bn, d, planes_P=1)          # dY planes of this conv's data / weight gradient (conv6: 'valid' -> padded by 1 too)
        else:
            d = ops.act_bwd(d, relu_out, ACT_RELU)
        d = engine.conv_bwd(ctx, s_c, d, need_dx=(idx > 0) or need_input_grad)
    if d is None:
        return None
    return d.view(B, 1, d.shape[1], d.shape[2])                          # NHWC with C == 1 is NCHW
