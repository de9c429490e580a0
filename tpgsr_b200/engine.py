"""Hand-sequenced forward/backward of the TPGSR building blocks on top of the C ABI (no autograd inside).

Every block is a pair ``*_fwd(ctx, ...) -> (out, saved)`` / ``*_bwd(ctx, saved, dout) -> din`` that enqueues kernels on the
current stream.  ``Ctx`` resolves reference-named parameters (``state_dict`` keys) to device tensors, hands out gradient
buffers (views of one flat buffer) and caches the packed weights for the lifetime of one forward(+backward).

Layout: activations are NHWC fp32 tensors [B,H,W,C]; ``x.view(-1, C)`` is the [M,C] pixel-row matrix the kernels see.
"""
from __future__ import annotations

import contextlib
from typing import Dict, Optional

import torch

from . import ops
from .ops import ACT_MISH, ACT_NONE, ACT_PRELU, ACT_RELU, ACT_TANH

Tensor = torch.Tensor


class Ctx:
    """Parameter / gradient resolver for one forward(+backward) pass of one module."""

    def __init__(self, params: Dict[str, Tensor], grads: Optional[Dict[str, Tensor]], training: bool,
                 need_input_grad: bool = False, share: Optional['Ctx'] = None, store: Optional['ops.PackStore'] = None):
        self.params = params          # name -> tensor (parameters and buffers, reference state_dict names + fused views)
        self.grads = grads            # name -> tensor (same names as params) or None when no backward is wanted
        self.training = training
        self.need_input_grad = need_input_grad
        # a second pass over the same module inside one step (shared SR model in the cascade) reuses the packed weights
        # and the first-write/accumulate bookkeeping of the first pass
        # (with a PackStore the packed weights outlive the step: the train step refreshes them after the optimiser, see ops.PackStore)
        self.store = share.store if share is not None else store
        self._packed: Dict[str, tuple] = share._packed if share is not None else (store.fp32 if store is not None else {})
        self._packed_bf16: Dict[str, tuple] = share._packed_bf16 if share is not None else (store.bf16 if store is not None else {})
        self._derived: Dict[str, tuple] = share._derived if share is not None else {}     # per-step operands computed FROM weights
        self._planes: Dict[int, tuple] = {}
        self._written = share._written if share is not None else set()
        # gradient side work (weight gradients, bias sums, GRU dW_hh) is queued on a second stream and runs concurrently with the
        # data-gradient chain; tensors it still reads are kept alive here until join() (the caching allocator reuses freed blocks)
        any_p = next(iter(params.values()))
        self.device = any_p.device
        self.side = ops.side_stream(self.device) if (grads is not None and any_p.is_cuda) else None
        self._keep = []

    def P(self, name: str) -> Tensor:
        return self.params[name]

    @contextlib.contextmanager
    def on_side(self, *keep):
        """Run the enclosed launches on the side stream, after everything queued so far on the current stream."""
        if self.side is None:
            yield
            return
        self._keep.extend(t for t in keep if t is not None)
        self.side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.side):
            yield

    def join(self) -> None:
        """The current stream waits for the side stream; side-stream operands may be released afterwards."""
        ops.join_side(self.device)
        self._keep.clear()

    def G(self, name: str):
        """(grad tensor, accumulate?) — first write in a backward overwrites, later writes accumulate."""
        acc = name in self._written or getattr(self, 'always_accumulate', False)
        self._written.add(name)
        return self.grads[name], acc

    def planes(self, x: Tensor):
        """bf16 hi/lo operand planes of an activation (cached: an activation feeds forward conv and weight-gradient)."""
        key = id(x)
        hit = self._planes.get(key)
        if hit is None:
            hit = (x, ops.split_bf16(x, ops.CONV_MODE == 'bf16x3'))
            self._planes[key] = hit
        return hit[1]

    def flat(self, x: Tensor, P: int, c_out: int = 0):
        """padded-flat bf16 planes of an activation (cached per (activation, padding, channel count))."""
        key = (id(x), 'flat', P, c_out or x.shape[-1])
        hit = self._planes.get(key)
        if hit is None:
            hit = (x, ops.split_bf16_flat(x, P, ops.CONV_MODE == 'bf16x3', c_out))
            self._planes[key] = hit
        return hit[1]

    def stacked(self, x: Tensor, KW: int, pw: int, P: int, sign: int):
        """kw-stacked padded-flat planes of a narrow activation (cached)."""
        key = (id(x), 'kwstack', KW, pw, P, sign)
        hit = self._planes.get(key)
        if hit is None:
            hit = (x, ops.kwstack_split(x, KW, pw, P, sign, ops.CONV_MODE == 'bf16x3'))
            self._planes[key] = hit
        return hit[1]

    def dflat(self, dy: Tensor, P: int, c_out: int = 0):
        """Operand planes of a GRADIENT tensor: flat(), or ONE bf16 plane when the gradient GEMMs run with two products (ops.BWD2)."""
        if not ops.BWD2:
            return self.flat(dy, P, c_out)
        key = (id(dy), 'dflat', P, c_out or dy.shape[-1])
        hit = self._planes.get(key)
        if hit is None:
            hit = (dy, ops.split_bf16_flat(dy, P, False, c_out))
            self._planes[key] = hit
        return hit[1]

    def dstacked(self, dy: Tensor, KW: int, pw: int, P: int, sign: int):
        """kw-stacked planes of a gradient tensor (see dflat)."""
        if not ops.BWD2:
            return self.stacked(dy, KW, pw, P, sign)
        key = (id(dy), 'dkwstack', KW, pw, P, sign)
        hit = self._planes.get(key)
        if hit is None:
            hit = (dy, ops.kwstack_split(dy, KW, pw, P, sign, False))
            self._planes[key] = hit
        return hit[1]

    def packed_kwstack(self, name: str, narrow_out: bool, transposed: bool, flip_kh: bool, flip_kw: bool):
        key = (name, 'kwstack', narrow_out, transposed, flip_kh, flip_kw)
        hit = self._packed_bf16.get(key)
        if hit is None:
            w = self.params[name]
            hit = ops.kwstack_pack(w, narrow_out, transposed, flip_kh, flip_kw, ops.CONV_MODE == 'bf16x3')
            self._packed_bf16[key] = hit
            if self.store is not None:
                Co, Ci, _, _ = w.shape
                narrow = (Ci if transposed else Co) if narrow_out else (Co if transposed else Ci)
                self.store.add('bf16', key, 2, w, hit, (narrow, narrow_out, transposed, int(flip_kh) | int(flip_kw) << 1, hit[0].shape[0]))
        return hit

    def planes_padded(self, x: Tensor, c_out: int):
        key = (id(x), c_out)
        hit = self._planes.get(key)
        if hit is None:
            hit = (x, ops.split_bf16_pad(x, c_out, ops.CONV_MODE == 'bf16x3'))
            self._planes[key] = hit
        return hit[1]

    def packed_bf16(self, name: str, as4d=None):
        hit = self._packed_bf16.get(name)
        if hit is None:
            w = self.params[name]
            if as4d is not None:
                w = w.view(*as4d)
            elif w.dim() == 2:
                w = w.view(w.shape[0], w.shape[1], 1, 1)
            hit = ops.pack_conv_weights_bf16(w, self.grads is not None, ops.CONV_MODE == 'bf16x3')
            self._packed_bf16[name] = hit
            if self.store is not None and '#' not in name:          # ('#': weights derived per step, e.g. InfoGen's equivalent convs)
                self.store.add('bf16', name, 0, w, hit[0] + hit[1])
        return hit

    def packed(self, name: str, need_dgrad: bool = True, as4d=None):
        key = name
        hit = self._packed.get(key)
        if hit is None or (need_dgrad and hit[1] is None):
            w = self.params[name]
            if as4d is not None:
                w = w.view(*as4d)
            elif w.dim() == 2:
                w = w.view(w.shape[0], w.shape[1], 1, 1)
            hit = ops.pack_conv_weights(w, need_dgrad)
            self._packed[key] = hit
            if self.store is not None:
                self.store.add('fp32', key, 1, w, hit)
        return hit


# ------------------------------------------------------------------------------------------------------------ conv
def _tc_ok(x: Tensor, Ci: int, Co: int, kh: int, kw: int, ph: int, pw: int, ldx: int) -> bool:
    B, H, W, Cx = x.shape
    return (ldx in (0, Cx) and Cx == Ci and 2 * ph == kh - 1 and 2 * pw == kw - 1 and x.is_contiguous()
            and ops.tc_supported(B, H, W, Ci, H, W, Co, kh, kw))


def _flat_ok(x: Tensor, Ci: int, Co: int, kh: int, kw: int, ph: int, pw: int, ldx: int) -> bool:
    """Stride-1 convs with 64 k input channels and any padding below the kernel size ('same', 'valid', 'full') run on the flat kernels."""
    B, H, W, Cx = x.shape
    return (ldx in (0, Cx) and Cx == Ci and 0 <= ph < kh and 0 <= pw < kw and x.is_contiguous()
            and ops.flat_supported(B, H, W, max(ph, pw), Ci, Co, kh, kw, ph, pw))


def _kwstack_kind(x: Tensor, Ci: int, Co: int, kh: int, kw: int, ph: int, pw: int, ldx: int, as4d):
    B, H, W, Cx = x.shape
    if as4d is not None or kh != kw or ph != pw or ldx not in (0, Cx) or Cx != Ci or not x.is_contiguous():
        return None
    return ops.kwstack_kind(B, H, W, Ci, Co, kh, ph)


def conv_fwd(ctx: Ctx, x: Tensor, wname: str, bname: Optional[str], k, pad, act=ACT_NONE, act_param: Optional[Tensor] = None,
             out_nchw: bool = False, need_dgrad: bool = True, as4d=None, ldx: int = 0, out: Optional[Tensor] = None,
             emit_planes: bool = False):
    """``emit_planes``: the caller knows the output feeds a 1x1 conv next; the flat epilogue then also writes its P = 0 bf16 planes."""
    w = ctx.P(wname)
    Co = w.shape[0]
    kh, kw = k if isinstance(k, tuple) else (k, k)
    ph, pw = pad if isinstance(pad, tuple) else (pad, pad)
    Ci = (w.numel() // Co) // (kh * kw)
    saved = (x, wname, bname, kh, kw, ph, pw, as4d, ldx)
    bias = ctx.P(bname) if bname else None
    kind = _kwstack_kind(x, Ci, Co, kh, kw, ph, pw, ldx, as4d) if out is None else None
    if kind == 'N':        # narrow output (final conv): 9x1 conv to the kw-stacked partial sums, then shift-and-add (+bias, act)
        wpk = ctx.packed_kwstack(wname, True, False, False, False)
        part = ops.conv_flat(ctx.flat(x, ph), wpk, None, kw * Co, kh, 1, ph, 0)
        return ops.kwsum(part, bias, Co, kw, pw, act, act_param, out_nchw), saved
    if kind == 'I':        # narrow input (block1): 9x1 conv on the kw-stacked input planes
        wpk = ctx.packed_kwstack(wname, False, False, False, False)
        y = ops.conv_flat(ctx.stacked(x, kw, pw, ph, 1), wpk, bias, Co, kh, 1, ph, 0, act, act_param, out_nchw)
        return y, saved
    if out is None and _flat_ok(x, Ci, Co, kh, kw, ph, pw, ldx):
        wf, _ = ctx.packed_bf16(wname, as4d)
        if emit_planes and not out_nchw and Co % 32 == 0:
            y, yp = ops.conv_flat(ctx.flat(x, max(ph, pw)), wf, bias, Co, kh, kw, ph, pw, act, act_param, emit_planes=True)
            ctx._planes[(id(y), 'flat', 0, Co)] = (y, yp)
            return y, saved
        y = ops.conv_flat(ctx.flat(x, max(ph, pw)), wf, bias, Co, kh, kw, ph, pw, act, act_param, out_nchw)
        return y, saved
    if out is None and _tc_ok(x, Ci, Co, kh, kw, ph, pw, ldx):
        B, H, W, _ = x.shape
        wf, _ = ctx.packed_bf16(wname, as4d)
        y = ops.conv_tc(ctx.planes(x), wf, bias, B, H, W, Ci, Co, kh, kw, ph, pw, act, act_param, out_nchw)
        return y, saved
    wf, _ = ctx.packed(wname, need_dgrad and ctx.grads is not None, as4d)
    y = ops.conv_fwd(x, wf, bias, Co, kh, kw, ph, pw, act, act_param, out_nchw, out=out, ldx=ldx)
    return y, saved


def conv_bwd(ctx: Ctx, saved, dy: Tensor, need_dx: bool = True, dx_out: Optional[Tensor] = None, dx_accumulate: bool = False,
             emit_dx_planes: bool = False):
    """dy: NHWC gradient w.r.t. the conv's pre-activation output.  ``emit_dx_planes``: dx feeds the backward of a 1x1 conv next."""
    x, wname, bname, kh, kw, ph, pw, as4d, ldx = saved
    w = ctx.P(wname)
    Co = w.shape[0]
    Ci = (w.numel() // Co) // (kh * kw)
    gw, acc = ctx.G(wname)
    B, H, W, Cx = x.shape
    P = max(ph, pw)
    kind = _kwstack_kind(x, Ci, Co, kh, kw, ph, pw, ldx, as4d) if dy.is_contiguous() else None
    if kind is not None:
        # operands shared by the weight gradient (side stream) and the data gradient (this stream) are produced here, before the fork
        if kind == 'N':     # final conv: dY is the narrow side -> stack it (sign -1), shared by the weight and the data gradient
            xp, dyp = ctx.flat(x, ph), ctx.dstacked(dy, kw, pw, ph, -1)
        else:               # block1: the stacked input planes of the forward pass against the plain dY planes
            xp, dyp = ctx.stacked(x, kw, pw, ph, 1), ctx.dflat(dy, ph)
        gb, accb = ctx.G(bname) if bname else (None, False)
        with ctx.on_side(dy):
            dwp = torch.empty(((kw * Co) if kind == 'N' else 64, 64, kh, 1), dtype=torch.float32, device=x.device)
            ops.conv_wgrad_flat(xp, dyp, dwp, (kw * Co) if kind == 'N' else 64, kh, 1, ph, 0, False)
            ops.kwstack_wgrad_scatter(dwp, gw, 0 if kind == 'N' else 1, acc)
            if bname:
                ops.colsum(dy.view(-1, Co), gb, accb)
        if not need_dx:
            return None
        if kind == 'N':     # dX[q][ci] = sum_t sum_j DYS[q + (t-ph)*Wp][j] * W[co(j)][ci][KH-1-t][kw(j)]
            wpk = ctx.packed_kwstack(wname, False, True, True, False)
            return ops.conv_flat(ctx.dstacked(dy, kw, pw, ph, -1), wpk, None, Ci, kh, 1, ph, 0, out=dx_out, accumulate=dx_accumulate)
        wpk = ctx.packed_kwstack(wname, True, True, True, True)      # narrow output: dX = shift-and-add of the stacked partial sums
        part = ops.conv_flat(ctx.dflat(dy, ph), wpk, None, kw * Ci, kh, 1, ph, 0)
        return ops.kwsum(part, None, Ci, kw, pw, out=dx_out, accumulate=dx_accumulate)
    same = (ldx in (0, Cx) and Cx == Ci and x.is_contiguous() and dy.is_contiguous() and 2 * ph == kh - 1 and 2 * pw == kw - 1)
    gb, accb = ctx.G(bname) if bname else (None, False)
    # flat weight gradient: x and dY planes on ONE plane grid.  'same' convs: both padded by P; 'valid' convs (CRNN conv6): dY
    # [B, H-KH+1, W-KW+1] padded up to x's grid
    Pdy = H + P - dy.shape[1]
    flat_w = (ldx in (0, Cx) and Cx == Ci and x.is_contiguous() and dy.is_contiguous() and 0 <= ph < kh and 0 <= pw < kw
              and Pdy == W + P - dy.shape[2] and Pdy >= 0 and ops.wgrad_flat_supported(B, H, W, P, Ci, Co, kh, kw))
    if flat_w:
        xp, dyp = ctx.flat(x, P), ctx.dflat(dy, Pdy, (Co + 63) // 64 * 64)         # shared with the data gradient: before the fork
        with ctx.on_side(dy):
            ops.conv_wgrad_flat(xp, dyp, gw, Co, kh, kw, ph, pw, acc)
            if bname:
                ops.colsum(dy.view(-1, Co), gb, accb)
    elif (same and ops.wgrad_tc_supported(B, H, W, Ci, Co, kh, kw)):
        xp = ctx.planes(x)
        dyp = ctx.planes(dy) if Co % 64 == 0 else ctx.planes_padded(dy, (Co + 63) // 64 * 64)
        with ctx.on_side(dy):
            ops.conv_wgrad_tc(xp, dyp, gw, B, H, W, Ci, Co, kh, kw, ph, pw, acc)
            if bname:
                ops.colsum(dy.view(-1, Co), gb, accb)
    else:
        with ctx.on_side(x, dy):
            ops.conv_wgrad(x, dy, gw, kh, kw, ph, pw, acc, ldx=ldx, Ci=Ci)
            if bname:
                ops.colsum(dy.view(-1, Co), gb, accb)
    if not need_dx:
        return None
    dense_out = dx_out is None or dx_out.shape[-1] == Ci
    Pd = max(kh - 1 - ph, kw - 1 - pw)                  # padding of the data-gradient conv = plane padding of dY ('same' convs: P)
    if dense_out and _flat_ok(dy, Co, Ci, kh, kw, kh - 1 - ph, kw - 1 - pw, 0):
        _, wd = ctx.packed_bf16(wname, as4d)
        if emit_dx_planes and dx_out is None and Ci % 32 == 0 and same:
            dx, dxp = ops.conv_flat(ctx.dflat(dy, Pd), wd, None, Ci, kh, kw, kh - 1 - ph, kw - 1 - pw, emit_planes=True, emit_lo=not ops.BWD2)
            ctx._planes[(id(dx), 'dflat' if ops.BWD2 else 'flat', 0, Ci)] = (dx, dxp)       # dx is the next backward conv's dY
            return dx
        return ops.conv_flat(ctx.dflat(dy, Pd), wd, None, Ci, kh, kw, kh - 1 - ph, kw - 1 - pw, out=dx_out, accumulate=dx_accumulate)
    if dense_out and _tc_ok(dy, Co, Ci, kh, kw, kh - 1 - ph, kw - 1 - pw, 0):
        _, wd = ctx.packed_bf16(wname, as4d)
        return ops.conv_tc(ctx.planes(dy), wd, None, B, H, W, Co, Ci, kh, kw, kh - 1 - ph, kw - 1 - pw, out=dx_out,
                           accumulate=dx_accumulate)
    _, wd = ctx.packed(wname, True, as4d)
    return ops.conv_dgrad(dy, wd, (x.shape[1], x.shape[2]), Ci, kh, kw, ph, pw, out=dx_out, accumulate=dx_accumulate)


# ------------------------------------------------------------------------------------------------------------ batch norm
def _planes_ok(x: Tensor) -> bool:
    return ops.BN_PLANES and ops.CONV_MODE != 'fp32' and x.dim() == 4 and x.shape[-1] % 64 == 0 and x.is_contiguous()


def bn_fwd(ctx: Ctx, x: Tensor, pre: str, act=ACT_NONE, residual: Optional[Tensor] = None, out: Optional[Tensor] = None,
           ldy: int = 0, planes_P: Optional[int] = None):
    """x [...,C] -> act(BN(x)) (+ residual).  Training: batch stats + running-stat update; eval: running stats.
    ``planes_P``: the result feeds a tensor-core conv with padding P next — the apply pass then also writes its bf16 operand planes."""
    C = x.shape[-1]
    x2 = x.view(-1, C)
    if ctx.training:
        mean_rstd, ss = ops.bn_stats(x2, ctx.P(pre + 'weight'), ctx.P(pre + 'bias'), ctx.P(pre + 'running_mean'),
                                     ctx.P(pre + 'running_var'))
    else:
        mean_rstd, ss = ops.bn_eval_prepare(ctx.P(pre + 'weight'), ctx.P(pre + 'bias'), ctx.P(pre + 'running_mean'),
                                            ctx.P(pre + 'running_var'))
    if planes_P is not None and out is None and _planes_ok(x):
        y, yp = ops.bn_apply_planes(x, ss, act, residual, planes_P, ops.CONV_MODE == 'bf16x3')
        ctx._planes[(id(y), 'flat', planes_P, C)] = (y, yp)
        return y, (x, pre, act, mean_rstd, ss, ctx.training)
    y = ops.bn_apply(x2, ss, act, residual.view(-1, C) if residual is not None else None,
                     out=out.view(-1, out.shape[-1]) if out is not None else None, ldy=ldy)
    if out is None:
        y = y.view(x.shape)
    else:
        y = out
    return y, (x, pre, act, mean_rstd, ss, ctx.training)


def bn_bwd(ctx: Ctx, saved, dy: Tensor, lddy: int = 0, planes_P: Optional[int] = None) -> Tensor:
    """``planes_P``: dx is the dY operand of a tensor-core conv backward with plane padding P — emitted by the apply pass."""
    x, pre, act, mean_rstd, ss, train = saved
    C = x.shape[-1]
    gg, acc = ctx.G(pre + 'weight')
    gb, _ = ctx.G(pre + 'bias')
    want_planes = planes_P is not None and _planes_ok(x)
    if (ops.BN_FUSED or want_planes) and lddy in (0, C) and x.dim() == 4 and C % 4 == 0 and x.is_contiguous() and dy.is_contiguous():
        need_lo = ops.CONV_MODE == 'bf16x3' and not ops.BWD2
        dx, dxp = ops.bn_bwd_fused(dy.view(x.shape), x, ss, mean_rstd, act, train, gg, gb, acc, planes_P if want_planes else None, need_lo)
        if dxp is not None:
            ctx._planes[(id(dx), 'dflat' if ops.BWD2 else 'flat', planes_P, C)] = (dx, dxp)
        return dx
    dx = ops.bn_bwd(dy.view(-1, dy.shape[-1]) if lddy == 0 else dy, x.view(-1, C), ss, mean_rstd, act, train, gg, gb, acc, lddy=lddy)
    return dx.view(x.shape)


# ------------------------------------------------------------------------------------------------------------ GruBlock
def _gru_proj_fused(ctx: Ctx, pre: str):
    """Packed bf16 operands of W_c = W_ih W_1x1 (and b_c = W_ih b_1x1 + b_ih), cached per step: ((fwd planes), (dgrad planes), b_c)."""
    key = pre + 'gru.fused'
    hit = ctx._derived.get(key)
    if hit is None:
        wc, bc = ops.fuse_linear(ctx.P(pre + 'gru.weight_ih'), ctx.P(pre + 'gru.bias_ih'), ctx.P(pre + 'conv1.weight'), ctx.P(pre + 'conv1.bias'))
        f, d = ops.pack_conv_weights_bf16(wc.view(wc.shape[0], wc.shape[1], 1, 1), ctx.grads is not None, ops.CONV_MODE == 'bf16x3')
        hit = (f, d, bc)
        ctx._derived[key] = hit
    return hit


def gru_block_fwd(ctx: Ctx, x: Tensor, pre: str, vertical: bool, residual: Optional[Tensor] = None, ldx: int = 0):
    """GruBlock (model/tsrn.py:491-508): 1x1 conv -> BiGRU along W (or along H when ``vertical``: the reference's
    transpose trick).  Returns (y or y+residual) as [B,H,W,64].

    The 1x1 conv and the GRU input projection are two linear maps with nothing in between (model/tsrn.py:499-504): when the input is a
    dense 64-channel map they run as ONE tensor-core GEMM with W_c = W_ih W_1x1; the backward pass recomputes the 1x1
    output it needs for dW_ih on the side stream."""
    B, H, W, Cx = x.shape
    whh, bhh = ctx.P(pre + 'gru.weight_hh'), ctx.P(pre + 'gru.bias_hh')
    res = residual.view(-1, 64) if residual is not None else None
    want = ops.FUSE_GRU_PROJ == '1' or (ops.FUSE_GRU_PROJ == 'infer' and ctx.grads is None)
    fused = (want and ldx in (0, Cx) and x.is_contiguous() and ops.flat_supported(B, H, W, 0, Cx, 192, 1, 1, 0, 0)
             and ops.flat_supported(B, H, W, 0, 192, Cx, 1, 1, 0, 0) and ops.wgrad_flat_supported(B, H, W, 0, 64, 192, 1, 1)
             and ops.wgrad_flat_supported(B, H, W, 0, Cx, 64, 1, 1))
    if fused:
        wf, _, bc = _gru_proj_fused(ctx, pre)
        gx = ops.conv_flat(ctx.flat(x, 0), wf, bc, 192, 1, 1, 0, 0)
        y, ysum, stash = ops.gru_fwd(gx.view(-1, 192), whh, bhh, B, H, W, vertical, res, need_stash=ctx.grads is not None)
        out = (ysum if residual is not None else y).view(B, H, W, 64)
        return out, ('fused', x, y, stash, pre, vertical, (B, H, W))
    y1, s_conv = conv_fwd(ctx, x, pre + 'conv1.weight', pre + 'conv1.bias', 1, 0, ldx=ldx, emit_planes=True)   # feeds the 1x1 projection
    gx, s_ih = conv_fwd(ctx, y1, pre + 'gru.weight_ih', pre + 'gru.bias_ih', 1, 0)
    y, ysum, stash = ops.gru_fwd(gx.view(-1, 192), whh, bhh, B, H, W, vertical, res, need_stash=ctx.grads is not None)
    out = (ysum if residual is not None else y).view(B, H, W, 64)
    return out, (s_conv, s_ih, y, stash, pre, vertical, (B, H, W))


def gru_block_bwd(ctx: Ctx, saved, dout: Tensor, need_dx: bool = True, dx_out: Optional[Tensor] = None,
                  dx_accumulate: bool = False):
    if saved[0] == 'fused':
        return _gru_block_bwd_fused(ctx, saved, dout, need_dx, dx_out, dx_accumulate)
    s_conv, s_ih, y, stash, pre, vertical, (B, H, W) = saved
    whh = ctx.P(pre + 'gru.weight_hh')
    gwhh, acc = ctx.G(pre + 'gru.weight_hh')
    gbhh, _ = ctx.G(pre + 'gru.bias_hh')
    gbih, _ = ctx.G(pre + 'gru.bias_ih')
    dgx, part = ops.gru_bwd_recurrence(dout.view(-1, 64), y, stash, whh, B, H, W, vertical)
    with ctx.on_side(dgx, y, stash, part):      # dW_hh outer products + bias reductions: off the data-gradient chain
        ops.gru_bwd_weights(dgx, y, stash, part, gwhh, gbhh, gbih, acc, B, H, W, vertical)
    dgx = dgx.view(B, H, W, 192)
    # input projection (bias grad already produced by the recurrence kernel)
    x_ih, wname, _, kh, kw, ph, pw, as4d, ldx = s_ih
    dy1 = conv_bwd(ctx, (x_ih, wname, None, kh, kw, ph, pw, as4d, ldx), dgx, emit_dx_planes=True)    # feeds the backward of the 1x1 conv
    return conv_bwd(ctx, s_conv, dy1, need_dx, dx_out, dx_accumulate)


def _gru_block_bwd_fused(ctx: Ctx, saved, dout: Tensor, need_dx: bool, dx_out: Optional[Tensor], dx_accumulate: bool):
    _, x, y, stash, pre, vertical, (B, H, W) = saved
    Cx = x.shape[-1]
    whh = ctx.P(pre + 'gru.weight_hh')
    gwhh, acc = ctx.G(pre + 'gru.weight_hh')
    gbhh, _ = ctx.G(pre + 'gru.bias_hh')
    gbih, _ = ctx.G(pre + 'gru.bias_ih')
    gwih, acc_ih = ctx.G(pre + 'gru.weight_ih')
    gw1, acc_w1 = ctx.G(pre + 'conv1.weight')
    gb1, acc_b1 = ctx.G(pre + 'conv1.bias')
    dgx, part = ops.gru_bwd_recurrence(dout.view(-1, 64), y, stash, whh, B, H, W, vertical)
    dgx4 = dgx.view(B, H, W, 192)
    xp, dgxp = ctx.flat(x, 0), ctx.flat(dgx4, 0)       # shared by both streams: produced before the fork
    with ctx.on_side(dgx, dgx4, y, stash, part, x):
        # everything that only feeds parameter gradients: dW_hh / biases, the recomputed 1x1 output, dW_ih, d(1x1 output), dW_1x1, db_1x1
        ops.gru_bwd_weights(dgx, y, stash, part, gwhh, gbhh, gbih, acc, B, H, W, vertical)
        w1f, _ = ctx.packed_bf16(pre + 'conv1.weight')
        y1 = ops.conv_flat(xp, w1f, ctx.P(pre + 'conv1.bias'), 64, 1, 1, 0, 0)
        ops.conv_wgrad_flat(ops.split_bf16_flat(y1, 0, ops.CONV_MODE == 'bf16x3'), dgxp, gwih, 192, 1, 1, 0, 0, acc_ih)
        _, wihd = ctx.packed_bf16(pre + 'gru.weight_ih')
        dy1 = ops.conv_flat(dgxp, wihd, None, 64, 1, 1, 0, 0)
        ops.conv_wgrad_flat(xp, ops.split_bf16_flat(dy1, 0, ops.CONV_MODE == 'bf16x3'), gw1, 64, 1, 1, 0, 0, acc_w1)
        ops.colsum(dy1.view(-1, 64), gb1, acc_b1)
    if not need_dx:
        return None
    _, wd, _ = _gru_proj_fused(ctx, pre)
    return ops.conv_flat(dgxp, wd, None, Cx, 1, 1, 0, 0, out=dx_out, accumulate=dx_accumulate)


# ------------------------------------------------------------------------------------------------------------ SRB
def srb_fwd(ctx: Ctx, x: Tensor, pre: str, text: Optional[Tensor] = None):
    """RecurrentResidualBlock / RecurrentResidualBlockTL (model/tsrn.py:373-426)."""
    B, H, W, C = x.shape
    c1, s_c1 = conv_fwd(ctx, x, pre + 'conv1.weight', pre + 'conv1.bias', 3, 1)
    a1, s_bn1 = bn_fwd(ctx, c1, pre + 'bn1.', ACT_MISH, planes_P=1)                 # feeds conv2 (3x3)
    c2, s_c2 = conv_fwd(ctx, a1, pre + 'conv2.weight', pre + 'conv2.bias', 3, 1)
    if text is None:
        r2, s_bn2 = bn_fwd(ctx, c2, pre + 'bn2.', planes_P=0)                       # feeds gru1's 1x1 conv
        cat = r2
    else:
        Ct = text.shape[-1]
        cat = ops.new((B, H, W, C + Ct), x)
        _, s_bn2 = bn_fwd(ctx, c2, pre + 'bn2.', out=cat, ldy=C + Ct)
        ops.copy_cols(text, Ct, cat, C + Ct, B * H * W, Ct, dst_off=C)
    r1sum, s_g1 = gru_block_fwd(ctx, cat, pre + 'gru1.', True, residual=x)     # x + gru1(...)
    out, s_g2 = gru_block_fwd(ctx, r1sum, pre + 'gru2.', False)
    return out, (s_c1, s_bn1, s_c2, s_bn2, s_g1, s_g2, text is not None, (B, H, W, C))


def srb_bwd(ctx: Ctx, saved, dout: Tensor, dtext: Optional[Tensor] = None):
    s_c1, s_bn1, s_c2, s_bn2, s_g1, s_g2, has_text, (B, H, W, C) = saved
    dsum = gru_block_bwd(ctx, s_g2, dout)                       # d(x + r1)
    dcat = gru_block_bwd(ctx, s_g1, dsum)                       # [B,H,W,64(+Ct)]
    if has_text:
        Ctot = dcat.shape[-1]
        Ct = Ctot - C
        ops.copy_cols(dcat, Ctot, dtext, Ct, B * H * W, Ct, accumulate=True, src_off=C)
        dc2 = bn_bwd(ctx, s_bn2, dcat.view(-1, Ctot), lddy=Ctot)
    else:
        dc2 = bn_bwd(ctx, s_bn2, dcat, planes_P=1)                                  # dY of conv2's gradients
    da1 = conv_bwd(ctx, s_c2, dc2)
    dc1 = bn_bwd(ctx, s_bn1, da1, planes_P=1)
    # dx = dgrad(conv1) + dsum : accumulate into the dsum buffer
    return conv_bwd(ctx, s_c1, dc1, True, dx_out=dsum, dx_accumulate=True)


# ------------------------------------------------------------------------------------------------------------ STN head
_STN_POOLS = {0: (2, 2), 2: (2, 2), 4: (2, 2), 6: (2, 2), 8: (1, 2)}


def stn_head_fwd(ctx: Ctx, x: Tensor, pre: str = 'stn_head.'):
    """STNHead.forward (model/stn_head.py:92-106).  x NHWC [B,16,64,4] -> (img_feat [B,512], ctrl [B,20,2])."""
    saved = []
    need_dgrad_first = False
    for i in (0, 2, 4, 6, 8, 10):
        q = f'{pre}stn_convnet.{i}.'
        c, s_c = conv_fwd(ctx, x, q + '0.weight', q + '0.bias', 3, 1, need_dgrad=(i != 0) or need_dgrad_first)
        a, s_bn = bn_fwd(ctx, c, q + '1.', ACT_RELU)
        if i in _STN_POOLS:
            k = _STN_POOLS[i]
            p, arg = ops.maxpool_fwd(a, k, k, need_argmax=ctx.grads is not None)
            saved.append((s_c, s_bn, (arg, tuple(a.shape), k)))
            x = p
        else:
            saved.append((s_c, s_bn, None))
            x = a
    B, Hh, Ww, Cc = x.shape       # [B,1,2,256]; NCHW flatten index c*Ww + w == fc1.weight viewed as [512, 256, 1, Ww]
    f_pre, s_fc1 = conv_fwd(ctx, x, pre + 'stn_fc1.0.weight', pre + 'stn_fc1.0.bias', (Hh, Ww), 0, as4d=(512, Cc, Hh, Ww))
    f, s_bn1 = bn_fwd(ctx, f_pre.view(B, 512), pre + 'stn_fc1.1.', ACT_RELU)
    f01 = ops.new((B, 1, 1, 512), x)
    ops.copy_cols(f, 512, f01, 512, B, 512, alpha=0.1)
    c, s_fc2 = conv_fwd(ctx, f01, pre + 'stn_fc2.weight', pre + 'stn_fc2.bias', 1, 0)
    nctrl = c.shape[-1] // 2
    return f, c.view(B, nctrl, 2), (saved, s_fc1, s_bn1, s_fc2, B)


def stn_head_bwd(ctx: Ctx, saved_all, dctrl: Tensor) -> None:
    saved, s_fc1, s_bn1, s_fc2, B = saved_all
    df01 = conv_bwd(ctx, s_fc2, dctrl.view(B, 1, 1, -1))
    df = ops.new((B, 512), dctrl)
    ops.copy_cols(df01, 512, df, 512, B, 512, alpha=0.1)
    dfpre = bn_bwd(ctx, s_bn1, df)
    d = conv_bwd(ctx, s_fc1, dfpre.view(B, 1, 1, 512))
    for idx in range(len(saved) - 1, -1, -1):
        s_c, s_bn, pool = saved[idx]
        if pool is not None:
            arg, in_shape, k = pool
            d = ops.maxpool_bwd(d, arg, in_shape, k, k)
        dc = bn_bwd(ctx, s_bn, d)
        d = conv_bwd(ctx, s_c, dc, need_dx=idx > 0)


# ------------------------------------------------------------------------------------------------------------ InfoGen
def info_gen_fwd(ctx: Ctx, prior: Tensor, pre: str, out_hw):
    raise NotImplementedError  # replaced in infogen.py once the transposed-conv kernels are registered


# ------------------------------------------------------------------------------------------------------------ TSRN
def tsrn_fwd(ctx: Ctx, x_nchw: Tensor, stn: bool, srb_nums: int = 5, text: Optional[Tensor] = None, text_saved=None):
    """TSRN.forward (model/tsrn.py:61-78) / TSRN_TL.forward (model/tsrn.py:178-215) after the prior has been expanded to
    ``text`` [B,16,64,32] (NHWC) by the InfoGen path.  Returns (sr NCHW [B,C,2H,2W], saved)."""
    x = ops.nchw_to_nhwc(x_nchw)
    B, H, W, C = x.shape
    s_stn = None
    if stn and ctx.training:
        _, ctrl, s_head = stn_head_fwd(ctx, x)
        xr, src = ops.tps_sample_fwd(x, ctrl, ctx.P('tps.inverse_kernel'), ctx.P('tps.reprT'))
        s_stn = (s_head, x, src)
        x_in = xr
    else:
        x_in = x
    b1pre, s_b1 = conv_fwd(ctx, x_in, 'block1.0.weight', 'block1.0.bias', 9, 4, need_dgrad=s_stn is not None)
    slope = ctx.P('block1.1.weight')
    b1 = ops.act_fwd(b1pre, ACT_PRELU, slope)
    b = b1
    s_srb = []
    for i in range(srb_nums):
        b, s = srb_fwd(ctx, b, f'block{i + 2}.', text)
        s_srb.append(s)
    k = srb_nums + 2
    c7, s_c7 = conv_fwd(ctx, b, f'block{k}.0.weight', f'block{k}.0.bias', 3, 1)
    s8, s_bn7 = bn_fwd(ctx, c7, f'block{k}.1.', ACT_NONE, residual=b1, planes_P=1)   # feeds the 64 -> 256 conv
    k += 1
    u, s_up = conv_fwd(ctx, s8, f'block{k}.0.conv.weight', f'block{k}.0.conv.bias', 3, 1)
    v = ops.pixelshuffle_mish_fwd(u)
    sr, s_out = conv_fwd(ctx, v, f'block{k}.1.weight', f'block{k}.1.bias', 9, 4, act=ACT_TANH, out_nchw=True)
    saved = (s_stn, s_b1, b1pre, s_srb, s_c7, s_bn7, s_up, u, s_out, sr, (B, H, W, C))
    return sr, saved


def tsrn_bwd(ctx: Ctx, saved, dsr_nchw: Tensor, dtext: Optional[Tensor] = None) -> None:
    """Backward of tsrn_fwd given d(loss)/d(sr) (NCHW, same strides as sr).  Parameter grads go to ctx.grads;
    ``dtext`` (NHWC [B,16,64,32], zero-initialised by the caller) accumulates the gradient of the expanded prior."""
    s_stn, s_b1, b1pre, s_srb, s_c7, s_bn7, s_up, u, s_out, sr, (B, H, W, C) = saved
    dpre = ops.nchw_to_nhwc(dsr_nchw, tanh_out=sr)              # tanh' fused into the layout change
    dv = conv_bwd(ctx, s_out, dpre)
    du = ops.pixelshuffle_mish_bwd(dv, u)
    ds8 = conv_bwd(ctx, s_up, du)                               # = d(b1 + bn7(c7)): flows to both operands
    dc7 = bn_bwd(ctx, s_bn7, ds8, planes_P=1)
    db = conv_bwd(ctx, s_c7, dc7)
    for s in reversed(s_srb):
        db = srb_bwd(ctx, s, db, dtext)
    db1 = ops.add(db, ds8)
    slope = ctx.P('block1.1.weight')
    gs, acc = ctx.G('block1.1.weight')
    db1pre = ops.act_bwd(db1, b1pre, ACT_PRELU, slope, gs, acc)
    if s_stn is None:
        conv_bwd(ctx, s_b1, db1pre, need_dx=False)
        return
    dxr = conv_bwd(ctx, s_b1, db1pre, need_dx=True)
    s_head, x, src = s_stn
    dctrl = ops.tps_sample_bwd(dxr, x, src, ctx.P('tps.inverse_kernel'), ctx.P('tps.reprT'))
    stn_head_bwd(ctx, s_head, dctrl)
