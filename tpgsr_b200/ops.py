"""Tensor-level wrappers over the C ABI (device pointers + dims on torch's current stream).

PyTorch is used here only as plumbing: device memory (``torch.empty``), streams and autograd hand-off.  All arithmetic
happens in libtpgsr_b200.so.  Activations are NHWC fp32 tensors of shape [B,H,W,C] (or [M,C]).
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch

from ._lib import LIB, call, query

ACT_NONE, ACT_RELU, ACT_PRELU, ACT_MISH, ACT_TANH = 0, 1, 2, 3, 4
OUT_NHWC, OUT_NCHW = 0, 1

# number of kernel launches issued by this process through the C ABI (bench.py reports it as `gpu_launches`)
LAUNCHES = 0
_KERNELS_PER_CALL = {
    'tpgsr_conv2d_wgrad': 2, 'tpgsr_bn_stats': 2, 'tpgsr_bn_bwd': 3, 'tpgsr_bn_bwd_fused': 2, 'tpgsr_colsum': 2, 'tpgsr_gru_bwd': 3,
    'tpgsr_image_loss_fwd': 2, 'tpgsr_semantic_loss_fwd': 2, 'tpgsr_optim_prepare': 2, 'tpgsr_lstm_bwd': 2,
    'tpgsr_conv2d_tc': 1, 'tpgsr_conv2d_wgrad_tc': 2, 'tpgsr_conv2d_flat': 1, 'tpgsr_conv2d_wgrad_flat': 2, 'tpgsr_gru_bwd_weights': 2,
    'tpgsr_psnr_ssim': 2,
}


def _call(name, *args):
    global LAUNCHES
    LAUNCHES += _KERNELS_PER_CALL.get(name, 1)
    call(name, *args)


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


_WS = {}
_WS_RETIRED = []     # outgrown buffers are kept alive: captured CUDA graphs may still reference their addresses


# ---- side stream: gradient side work (weight gradients, bias column sums, GRU dW_hh) runs concurrently with the data-gradient chain ------
# TPGSR_SIDE_STREAM=0 keeps everything on one stream.
SIDE_STREAM = os.environ.get('TPGSR_SIDE_STREAM', '1') != '0'
_SIDE = {}


def _dev_index(device) -> int:
    d = torch.device(device)
    return d.index if d.index is not None else torch.cuda.current_device()


def side_stream(device):
    """The per-device side stream (None when disabled)."""
    if not SIDE_STREAM:
        return None
    key = _dev_index(device)
    st = _SIDE.get(key)
    if st is None:
        st = torch.cuda.Stream(device=torch.device('cuda', key))
        _SIDE[key] = st
    return st


def join_side(device) -> None:
    """Make the current stream wait for everything queued on the device's side stream (no-op if nothing was ever forked)."""
    st = _SIDE.get(_dev_index(device))
    if st is not None:
        cur = torch.cuda.current_stream(torch.device('cuda', _dev_index(device)))
        if cur.cuda_stream != st.cuda_stream:
            cur.wait_stream(st)


class on_stream_after_current:
    """``with on_stream_after_current(side, device):`` — run the enclosed launches on ``side`` after everything queued so far on the
    current stream (no-op when ``side`` is None).  The caller joins with :func:`join_side` before consuming the results."""

    def __init__(self, side, device):
        self.side, self.device, self._ctx = side, device, None

    def __enter__(self):
        if self.side is not None:
            self.side.wait_stream(torch.cuda.current_stream(torch.device('cuda', _dev_index(self.device))))
            self._ctx = torch.cuda.stream(self.side)
            self._ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self._ctx is not None:
            self._ctx.__exit__(*exc)
        return False


def workspace(nbytes: int, device) -> Tuple[int, int]:
    """A scratch buffer per (device, main / side stream), grown on demand; stable pointer once large enough — safe under graph
    capture because every kernel that uses one buffer runs on the same stream."""
    idx = _dev_index(device)
    side = _SIDE.get(idx)
    on_side = side is not None and torch.cuda.current_stream(torch.device('cuda', idx)).cuda_stream == side.cuda_stream
    key = (idx, on_side)
    buf = _WS.get(key)
    need = max(int(nbytes), 1)
    if buf is None or buf.numel() < need:
        if buf is not None:
            _WS_RETIRED.append(buf)
        buf = torch.empty(max(need, 64 << 20), dtype=torch.uint8, device=device)
        _WS[key] = buf
    return buf.data_ptr(), buf.numel()


def prime_workspaces(nbytes: int, device) -> None:
    """Allocate the main and side scratch buffers up front (outside any graph capture)."""
    workspace(nbytes, device)
    st = side_stream(device)
    if st is not None:
        with torch.cuda.stream(st):
            workspace(nbytes, device)
        torch.cuda.current_stream(torch.device('cuda', _dev_index(device))).wait_stream(st)


def new(shape, like: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    return torch.empty(shape, dtype=dtype, device=like.device)


# ------------------------------------------------------------------------------------------------ layout
def nchw_to_nhwc(x: torch.Tensor, tanh_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Strided NCHW view -> dense NHWC [B,H,W,C]; with ``tanh_out`` multiplies by (1-tanh_out^2) (tanh backward)."""
    B, C, H, W = x.shape
    if tanh_out is not None:
        assert tanh_out.stride() == x.stride()
    y = new((B, H, W, C), x)
    sb, sc, sh, sw = x.stride()
    _call('tpgsr_nchw_to_nhwc', ptr(x), ptr(y), B, C, H, W, sb, sc, sh, sw, C, ptr(tanh_out), stream())
    return y


def upsample_nearest_to_nhwc(x: torch.Tensor, scale: int) -> torch.Tensor:
    """Strided NCHW view -> dense NHWC [B, H*scale, W*scale, C] with nearest-neighbour replication."""
    B, C, H, W = x.shape
    y = new((B, H * scale, W * scale, C), x)
    sb, sc, sh, sw = x.stride()
    _call('tpgsr_upsample_nearest_nchw_to_nhwc', ptr(x), ptr(y), B, C, H, W, scale, sb, sc, sh, sw, stream())
    return y


def nhwc_to_nchw(x: torch.Tensor) -> torch.Tensor:
    B, H, W, C = x.shape
    y = new((B, C, H, W), x)
    _call('tpgsr_nhwc_to_nchw', ptr(x), ptr(y), B, C, H, W, C, stream())
    return y


# ------------------------------------------------------------------------------------------------ conv / linear
def pack_conv_weights(w: torch.Tensor, need_dgrad: bool = True):
    Co, Ci, KH, KW = w.shape
    wf = new((KH * KW * Ci, Co), w)
    wd = new((KH * KW * Co, Ci), w) if need_dgrad else None
    _call('tpgsr_pack_conv_weights', ptr(w), ptr(wf), ptr(wd), Co, Ci, KH, KW, stream())
    return wf, wd


def conv_fwd(x: torch.Tensor, wk_fwd: torch.Tensor, bias: Optional[torch.Tensor], Co: int, KH: int, KW: int, ph: int, pw: int,
             act: int = ACT_NONE, act_param: Optional[torch.Tensor] = None, out_nchw: bool = False,
             out: Optional[torch.Tensor] = None, ldx: int = 0) -> torch.Tensor:
    """y = act(conv(x) + bias); x NHWC [B,Hi,Wi,Ci(pitch ldx)] -> NHWC [B,Ho,Wo,Co] (or NCHW)."""
    B, Hi, Wi, Cx = x.shape
    Ci = wk_fwd.shape[0] // (KH * KW)
    Ho, Wo = Hi + 2 * ph - KH + 1, Wi + 2 * pw - KW + 1
    if out is None:
        out = new((B, Co, Ho, Wo) if out_nchw else (B, Ho, Wo, Co), x)
    wp, wn = workspace(query('tpgsr_conv2d_igemm_workspace_bytes', B, Hi, Wi, Ci, Ho, Wo, Co, KH, KW), x.device)
    _call('tpgsr_conv2d_igemm_ws', ptr(x), ptr(wk_fwd), ptr(bias), ptr(out), B, Hi, Wi, Ci, Ho, Wo, Co, KH, KW, ph, pw,
          ldx or Cx, Co, act, ptr(act_param), OUT_NCHW if out_nchw else OUT_NHWC, 0, wp, wn, stream())
    return out


def conv_dgrad(dy: torch.Tensor, wk_dgrad: torch.Tensor, in_hw: Tuple[int, int], Ci: int, KH: int, KW: int, ph: int, pw: int,
               out: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    """dx = conv_transpose(dy): gather-GEMM over dy with the flipped packing.  dy NHWC [B,Ho,Wo,Co] -> [B,Hi,Wi,Ci]."""
    B, Ho, Wo, Co = dy.shape
    Hi, Wi = in_hw
    if out is None:
        out = new((B, Hi, Wi, Ci), dy)
        accumulate = False
    wp, wn = workspace(query('tpgsr_conv2d_igemm_workspace_bytes', B, Ho, Wo, Co, Hi, Wi, Ci, KH, KW), dy.device)
    _call('tpgsr_conv2d_igemm_ws', ptr(dy), ptr(wk_dgrad), None, ptr(out), B, Ho, Wo, Co, Hi, Wi, Ci, KH, KW,
          KH - 1 - ph, KW - 1 - pw, Co, out.shape[-1], ACT_NONE, None, OUT_NHWC, int(accumulate), wp, wn, stream())
    return out


def conv_wgrad(x: torch.Tensor, dy: torch.Tensor, dw: torch.Tensor, KH: int, KW: int, ph: int, pw: int,
               accumulate: bool, ldx: int = 0, Ci: int = 0) -> None:
    B, Hi, Wi, Cx = x.shape
    _, Ho, Wo, Co = dy.shape
    Ci = Ci or Cx
    nbytes = query('tpgsr_conv2d_wgrad_workspace_bytes', B, Hi, Wi, Ci, Ho, Wo, Co, KH, KW)
    wp, wn = workspace(nbytes, x.device)
    _call('tpgsr_conv2d_wgrad', ptr(x), ptr(dy), ptr(dw), B, Hi, Wi, Ci, Ho, Wo, Co, KH, KW, ph, pw, ldx or Cx, Co,
          int(accumulate), wp, wn, stream())


# ---- tensor-core path (tcgen05): bf16 hi/lo operand planes -------------------------------------------------------------
# 'bf16x3' = split-bf16, three MMAs per K step, fp32-class accuracy (the parity mode); 'bf16' = one MMA (throughput mode,
# misses the 1e-3 parity budget — every number measured in it is labelled); 'fp32' = CUDA-core path only.
CONV_MODE = os.environ.get('TPGSR_CONV_MODE', 'bf16x3')


def tc_supported(B, Hi, Wi, Ci, Ho, Wo, Co, KH, KW) -> bool:
    return CONV_MODE != 'fp32' and bool(LIB.tpgsr_conv2d_tc_supported(B, Hi, Wi, Ci, Ho, Wo, Co, KH, KW))


def split_bf16(x: torch.Tensor, need_lo: bool = True):
    """fp32 NHWC [..., C] (dense) -> (hi, lo) bf16 planes of the same shape; lo is None in 'bf16' mode."""
    C = x.shape[-1]
    M = x.numel() // C
    hi = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    lo = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device) if need_lo else None
    _call('tpgsr_split_bf16', ptr(x), C, ptr(hi), ptr(lo), M, C, stream())
    return hi, lo


def pack_conv_weights_bf16(w: torch.Tensor, need_dgrad: bool, need_lo: bool = True):
    Co, Ci, KH, KW = w.shape
    mk = lambda r, c: torch.empty((r, c), dtype=torch.bfloat16, device=w.device)
    f_hi = mk(Co, KH * KW * Ci)
    f_lo = mk(Co, KH * KW * Ci) if need_lo else None
    d_hi = mk(Ci, KH * KW * Co) if need_dgrad else None
    d_lo = mk(Ci, KH * KW * Co) if (need_dgrad and need_lo) else None
    _call('tpgsr_pack_conv_weights_bf16', ptr(w), ptr(f_hi), ptr(f_lo), ptr(d_hi), ptr(d_lo), Co, Ci, KH, KW, stream())
    return (f_hi, f_lo), (d_hi, d_lo)


def conv_tc(x_planes, w_planes, bias, B, Hi, Wi, Ci, Co, KH, KW, ph, pw, act=ACT_NONE, act_param=None, out_nchw=False,
            out: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    """Tensor-core conv on pre-split operands ('same' geometry: Ho=Hi, Wo=Wi)."""
    hi, lo = x_planes
    if out is None:
        out = torch.empty((B, Co, Hi, Wi) if out_nchw else (B, Hi, Wi, Co), dtype=torch.float32, device=hi.device)
        accumulate = False
    nprod = 3 if (CONV_MODE == 'bf16x3' and lo is not None and w_planes[1] is not None) else 1
    _call('tpgsr_conv2d_tc', ptr(hi), ptr(lo), ptr(w_planes[0]), ptr(w_planes[1]), ptr(bias), ptr(out), B, Hi, Wi, Ci, Hi, Wi,
          Co, KH, KW, ph, pw, out.shape[-1] if not out_nchw else Co, act, ptr(act_param), OUT_NCHW if out_nchw else OUT_NHWC,
          int(accumulate), nprod, stream())
    return out


def split_bf16_pad(x: torch.Tensor, c_out: int, need_lo: bool = True):
    """fp32 NHWC [..., C] -> bf16 planes [..., c_out] with zero channels beyond C (ragged Co for the wgrad B' operand)."""
    C = x.shape[-1]
    M = x.numel() // C
    shape = tuple(x.shape[:-1]) + (c_out,)
    hi = torch.empty(shape, dtype=torch.bfloat16, device=x.device)
    lo = torch.empty(shape, dtype=torch.bfloat16, device=x.device) if need_lo else None
    _call('tpgsr_split_bf16_pad', ptr(x), C, ptr(hi), ptr(lo), M, C, c_out, stream())
    return hi, lo


def wgrad_tc_supported(B, H, W, Ci, Co, KH, KW) -> bool:
    if CONV_MODE == 'fp32':
        return False
    cpad = (Co + 63) // 64 * 64
    return bool(LIB.tpgsr_conv2d_wgrad_tc_supported(B, H, W, Ci, Co, cpad, KH, KW, 3 if CONV_MODE == 'bf16x3' else 1))


def conv_wgrad_tc(x_planes, dy_planes, dw: torch.Tensor, B, H, W, Ci, Co, KH, KW, ph, pw, accumulate: bool) -> None:
    cpad = dy_planes[0].shape[-1]
    nprod = 3 if (CONV_MODE == 'bf16x3' and x_planes[1] is not None and dy_planes[1] is not None) else 1
    nbytes = query('tpgsr_conv2d_wgrad_tc_workspace_bytes', B, H, W, Ci, Co, cpad, KH, KW, nprod)
    wp, wn = workspace(nbytes, dw.device)
    _call('tpgsr_conv2d_wgrad_tc', ptr(x_planes[0]), ptr(x_planes[1]), ptr(dy_planes[0]), ptr(dy_planes[1]), ptr(dw), B, H, W, Ci,
          Co, cpad, KH, KW, ph, pw, int(accumulate), nprod, wp, wn, stream())


# ---- padded-flat tensor-core path (conv_flat.cu): operand planes [B][H+P][W+P][C] with zero pad rows / columns ------------------
# TPGSR_FLAT=0 falls back to the per-tap TMA kernels above (kept for the shapes the flat kernels do not take).
FLAT = os.environ.get('TPGSR_FLAT', '1') != '0'
# GruBlock 1x1 conv + GRU input projection as one GEMM (W_c = W_ih W_1x1): 'infer' (default) = forward-only passes, '1' = training too,
# '0' = never.  Measured on B200 (round 1): in training the fused path is neutral (5.77 vs 5.70 ms/step) — the backward pass has to
# recompute the 1x1 output for dW_ih, which costs what the forward pass saves — so it is kept for inference only.
FUSE_GRU_PROJ = os.environ.get('TPGSR_FUSE_GRU_PROJ', 'infer')

# developer A/B switch: TPGSR_FLAT_TALL=0 disables the tall plan of the flat convs (KH x 1 kernels: the kw-stacked 9x9 convs)
FLAT_TALL = os.environ.get('TPGSR_FLAT_TALL', '1') != '0'
# developer A/B switch: TPGSR_FLAT_WS=0 disables the weights-stationary plan (weights resident in tensor memory) of the 3x3 64 -> 64 convs
FLAT_WS = os.environ.get('TPGSR_FLAT_WS', '1') != '0'
if not (FLAT_TALL and FLAT_WS):
    LIB.tpgsr_conv2d_flat_set_debug_skip((0 if FLAT_TALL else 8) | (0 if FLAT_WS else 16))


class FlatPlanes:
    """bf16 hi/lo planes of an NHWC activation in the padded-flat layout."""
    __slots__ = ('hi', 'lo', 'B', 'H', 'W', 'C', 'P')

    def __init__(self, hi, lo, B, H, W, C, P):
        self.hi, self.lo, self.B, self.H, self.W, self.C, self.P = hi, lo, B, H, W, C, P


def nprod() -> int:
    return 3 if CONV_MODE == 'bf16x3' else 1


# OPT-IN (default off: the parity mode keeps three products everywhere): gradient GEMMs with TWO products — dY enters as ONE bf16 plane (no lo
# plane is split, staged or multiplied) against the exact (hi + lo) weights / activations: dX = dY_hi (W_hi + W_lo), dW = (X_hi + X_lo) dY_hi.
# The forward pass keeps three products, so the SR output is untouched; on the oracle graph the non-STN gradients move from 1.7e-4 to 2.9e-4
# relative L2 against fp32 (scripts/study_operand_precision.py, mode 'bwd2').  Measured on B200 (config 3, B = 48): 5 156 against 4 832 images/s
# (+6.7%), with TPGSR_WGRAD_PRODUCTS=1 on top 5 283 (+9.3%) — profiles/r02/bench_c3_bwd2*_call5.json; tests/test_gpu_backward_products.py.
BWD2 = os.environ.get('TPGSR_BWD_PRODUCTS', '3') == '2' and CONV_MODE == 'bf16x3'
# OPT-IN (default off): the flat weight gradient with ONE product (X_hi dY_hi, the kernel's bf16 instantiation).  A weight gradient
# sums over every pixel of the batch (49 152 at B = 48), so the operands' independent rounding errors average out: on the oracle graph the
# non-STN gradients are at 3.1e-4 relative L2 with it (2.9e-4 with two products, 1.7e-4 with three; study script, mode 'bwd2w1').  Off by default.
WGRAD1 = os.environ.get('TPGSR_WGRAD_PRODUCTS', '3') == '1' and CONV_MODE == 'bf16x3'


def _nprod_for(a_lo, b_lo) -> int:
    """Products the flat kernels run for the planes at hand: 3 = both operands split, 2 = only the B / X side split, 1 = bf16."""
    if nprod() != 3 or b_lo is None:
        return 1
    return 3 if a_lo is not None else 2


def split_bf16_flat(x: torch.Tensor, P: int, need_lo: bool = True, c_out: int = 0) -> FlatPlanes:
    """fp32 NHWC [B,H,W,C] (dense) -> padded-flat bf16 planes with c_out >= C channels (extra channels zero)."""
    B, H, W, C = x.shape
    c_out = c_out or C
    rows = B * (H + P) * (W + P)
    hi = torch.empty((rows, c_out), dtype=torch.bfloat16, device=x.device)
    lo = torch.empty((rows, c_out), dtype=torch.bfloat16, device=x.device) if need_lo else None
    _call('tpgsr_split_bf16_flat', ptr(x), C, ptr(hi), ptr(lo), B, H, W, C, c_out, P, stream())
    return FlatPlanes(hi, lo, B, H, W, c_out, P)


def flat_supported(B, H, W, P, Ci, Co, KH, KW, ph, pw) -> bool:
    return (FLAT and CONV_MODE != 'fp32'
            and bool(LIB.tpgsr_conv2d_flat_supported(B, H, W, P, Ci, Co, KH, KW, ph, pw, nprod())))


def conv_flat(xp: FlatPlanes, w_planes, bias, Co, KH, KW, ph, pw, act=ACT_NONE, act_param=None, out_nchw=False,
              out: Optional[torch.Tensor] = None, accumulate: bool = False, emit_planes: bool = False, emit_lo: bool = True):
    """'same' conv on padded-flat planes -> dense fp32 NHWC [B,H,W,Co] (or NCHW).  ``emit_planes``: also return the output's dense
    bf16 hi/lo planes (= its P = 0 operand planes) written by the same epilogue -> (out, FlatPlanes)."""
    B, H, W = xp.B, xp.H, xp.W
    Ho, Wo = H + 2 * ph - KH + 1, W + 2 * pw - KW + 1           # 'same' convs: (H, W); 'valid': smaller; 'full': larger (fits the plane grid)
    if out is None:
        out = torch.empty((B, Co, Ho, Wo) if out_nchw else (B, Ho, Wo, Co), dtype=torch.float32, device=xp.hi.device)
        accumulate = False
    n = _nprod_for(xp.lo, w_planes[1])
    if emit_planes:
        assert not out_nchw and Co % 32 == 0 and out.shape[-1] == Co and (Ho, Wo) == (H, W)
        hi = torch.empty((B * H * W, Co), dtype=torch.bfloat16, device=out.device)
        lo = torch.empty((B * H * W, Co), dtype=torch.bfloat16, device=out.device) if (nprod() == 3 and emit_lo) else None
        _call('tpgsr_conv2d_flat_planes', ptr(xp.hi), ptr(xp.lo), ptr(w_planes[0]), ptr(w_planes[1]), ptr(bias), ptr(out), ptr(hi), ptr(lo),
              B, H, W, xp.P, xp.C, Co, KH, KW, ph, pw, Co, act, ptr(act_param), OUT_NHWC, int(accumulate), n, stream())
        return out, FlatPlanes(hi, lo, B, H, W, Co, 0)
    _call('tpgsr_conv2d_flat', ptr(xp.hi), ptr(xp.lo), ptr(w_planes[0]), ptr(w_planes[1]), ptr(bias), ptr(out), B, H, W, xp.P,
          xp.C, Co, KH, KW, ph, pw, out.shape[-1] if not out_nchw else Co, act, ptr(act_param),
          OUT_NCHW if out_nchw else OUT_NHWC, int(accumulate), n, stream())
    return out


def wgrad_flat_supported(B, H, W, P, Ci, Co, KH, KW) -> bool:
    cpad = (Co + 63) // 64 * 64
    return (FLAT and CONV_MODE != 'fp32'
            and bool(LIB.tpgsr_conv2d_wgrad_flat_supported(B, H, W, P, Ci, Co, cpad, KH, KW, nprod())))


def conv_wgrad_flat(xp: FlatPlanes, dyp: FlatPlanes, dw: torch.Tensor, Co, KH, KW, ph, pw, accumulate: bool) -> None:
    assert (xp.B, xp.H + xp.P, xp.W + xp.P) == (dyp.B, dyp.H + dyp.P, dyp.W + dyp.P)      # same plane grid (valid convs: dY is padded up to it)
    n = 1 if WGRAD1 else _nprod_for(dyp.lo, xp.lo)            # 2: dY one plane against X hi + lo
    nbytes = query('tpgsr_conv2d_wgrad_flat_workspace_bytes', xp.B, xp.H, xp.W, xp.P, xp.C, Co, dyp.C, KH, KW, n)
    wp, wn = workspace(nbytes, dw.device)
    _call('tpgsr_conv2d_wgrad_flat', ptr(xp.hi), ptr(xp.lo), ptr(dyp.hi), ptr(dyp.lo), ptr(dw), xp.B, xp.H, xp.W, xp.P, xp.C, Co,
          dyp.C, KH, KW, ph, pw, int(accumulate), n, wp, wn, stream())


# ---- kw-stacking (kwstack.cu): 9x9 convs with a narrow (<= 64/9 channel) side as 9x1 tensor-core convs -----------------------------------
KWSTACK = os.environ.get('TPGSR_KWSTACK', '1') != '0'


def kwstack_kind(B, H, W, Ci, Co, k, pad):
    """'N' (narrow output: Ci == 64, Co*k <= 64), 'I' (narrow input: Co == 64, Ci*k <= 64) or None."""
    if not (FLAT and KWSTACK) or CONV_MODE == 'fp32' or k < 5 or 2 * pad != k - 1:
        return None
    if Ci == 64 and Co * k <= 64:
        kind, co_eff = 'N', Co * k
    elif Co == 64 and Ci * k <= 64:
        kind, co_eff = 'I', 64
    else:
        return None
    if not LIB.tpgsr_conv2d_flat_supported(B, H, W, pad, 64, co_eff, k, 1, pad, 0, nprod()):
        return None
    if not LIB.tpgsr_conv2d_wgrad_flat_supported(B, H, W, pad, 64, co_eff, 64, k, 1, nprod()):
        return None
    return kind


def kwstack_pack(w: torch.Tensor, narrow_out: bool, transposed: bool, flip_kh: bool, flip_kw: bool, need_lo: bool = True):
    """OIHW fp32 -> K-major bf16 hi/lo [rows][KH*64] for the KHx1 stacked conv (see kwstack.cu)."""
    Co, Ci, KH, KW = w.shape
    narrow = (Ci if transposed else Co) if narrow_out else (Co if transposed else Ci)
    rows = (KW * narrow + 15) // 16 * 16 if narrow_out else 64
    hi = torch.empty((rows, KH * 64), dtype=torch.bfloat16, device=w.device)
    lo = torch.empty((rows, KH * 64), dtype=torch.bfloat16, device=w.device) if need_lo else None
    _call('tpgsr_kwstack_pack_weights', ptr(w), ptr(hi), ptr(lo), Co, Ci, KH, KW, int(narrow_out), int(transposed), int(flip_kh),
          int(flip_kw), rows, stream())
    return hi, lo


# ---- persistent packed weights (train steps): filled by the per-weight kernels above during the first step, then refreshed after every
# optimiser step by ONE launch of tpgsr_pack_weights_batched (pack_batched.cu) ----
PREPACK = os.environ.get('TPGSR_PREPACK', '0') != '0'


class PackStore:
    """Packed operand copies of a module's weights that outlive one ``engine.Ctx``.

    ``bf16`` / ``fp32`` are the dicts a Ctx caches its packs in; with a store they are shared by every Ctx of the train step, so only the
    first step launches the per-weight pack kernels.  Each miss also registers a job (source view, outputs, geometry); ``refresh()`` re-packs
    all registered weights in one launch — the train step calls it after Adam (the packed copies follow the weights) and whenever the
    parameters are changed from outside."""

    def __init__(self):
        self.bf16 = {}
        self.fp32 = {}
        self.jobs = {}          # (dict id, key) -> (kind, w, outs, geometry)
        self._table = None
        self._n = 0
        self._max = 0

    def add(self, which: str, key, kind: int, w: torch.Tensor, outs, extra=(0, 0, 0, 0, 0)) -> None:
        self.jobs[(which, key)] = (kind, w, tuple(outs), tuple(int(e) for e in extra))
        self._table = None

    def _build(self) -> None:
        import struct
        size = LIB.tpgsr_pack_job_bytes()
        fmt = '5Q10i'
        if struct.calcsize(fmt) != size:
            raise RuntimeError(f'PackJob layout mismatch: library says {size} bytes, host packs {struct.calcsize(fmt)}')
        blob = bytearray()
        dev, mx = None, 1
        for kind, w, outs, extra in self.jobs.values():
            Co, Ci, KH, KW = w.shape
            if not w.is_contiguous():
                raise RuntimeError('PackStore: weight views must be contiguous')
            o = [ptr(t) or 0 for t in outs] + [0] * (4 - len(outs))
            blob += struct.pack(fmt, w.data_ptr(), *o, kind, Co, Ci, KH, KW, *extra)
            mx = max(mx, (extra[4] * KH * 64) if kind == 2 else Co * Ci * KH * KW)
            dev = w.device
        self._n, self._max = len(self.jobs), mx
        self._table = torch.frombuffer(blob, dtype=torch.uint8).clone().to(dev) if self._n else None

    def refresh(self) -> None:
        """Re-pack every registered weight from its current values (one launch on the current stream; graph-capturable)."""
        if not self.jobs:
            return
        if self._table is None:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError('PackStore: the job table must be built before graph capture (run one eager step first)')
            self._build()
        _call('tpgsr_pack_weights_batched', ptr(self._table), self._n, self._max, stream())


def kwstack_split(x: torch.Tensor, KW: int, pw: int, P: int, sign: int, need_lo: bool = True) -> FlatPlanes:
    """fp32 NHWC [B,H,W,C] (C*KW <= 64) -> stacked padded-flat planes [B][H+P][W+P][64], channel kw*C+i = x[.., c + sign*(kw-pw), i]."""
    B, H, W, C = x.shape
    rows = B * (H + P) * (W + P)
    hi = torch.empty((rows, 64), dtype=torch.bfloat16, device=x.device)
    lo = torch.empty((rows, 64), dtype=torch.bfloat16, device=x.device) if need_lo else None
    _call('tpgsr_kwstack_split', ptr(x), C, ptr(hi), ptr(lo), B, H, W, C, KW, pw, P, sign, stream())
    return FlatPlanes(hi, lo, B, H, W, 64, P)


def kwsum(p: torch.Tensor, bias, C: int, KW: int, pw: int, act=ACT_NONE, act_param=None, out_nchw=False,
          out: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    """y[.., c, o] = act(bias[o] + sum_kw p[.., c + kw - pw, kw*C + o]);  p NHWC [B,H,W,>=KW*C] -> NHWC [B,H,W,C] or NCHW."""
    B, H, W, ldp = p.shape
    if out is None:
        out = torch.empty((B, C, H, W) if out_nchw else (B, H, W, C), dtype=torch.float32, device=p.device)
        accumulate = False
    _call('tpgsr_kwsum', ptr(p), ldp, ptr(bias), ptr(out), B, H, W, C, KW, pw, C if out_nchw else out.shape[-1], act, ptr(act_param),
          OUT_NCHW if out_nchw else OUT_NHWC, int(accumulate), stream())
    return out


def kwstack_wgrad_scatter(dwp: torch.Tensor, dw: torch.Tensor, layout: int, accumulate: bool) -> None:
    Co, Ci, KH, KW = dw.shape
    _call('tpgsr_kwstack_wgrad_scatter', ptr(dwp), ptr(dw), Co, Ci, KH, KW, layout, int(accumulate), stream())


def colsum(x2d: torch.Tensor, out: torch.Tensor, accumulate: bool) -> None:
    """out[C] (+)= sum over rows of x [M,C] (bias gradients)."""
    M, C = x2d.shape
    wp, wn = workspace(query('tpgsr_colreduce_workspace_bytes', M, C), x2d.device)
    _call('tpgsr_colsum', ptr(x2d), M, C, C, ptr(out), int(accumulate), wp, wn, stream())


# ------------------------------------------------------------------------------------------------ batch norm
_TICKETS = {}


def ticket(device) -> int:
    """One zeroed unsigned per (device, main / side stream) for the "last block finalizes" reductions: kernels that use it run on one
    stream (ordered) and leave it zero, so the pointer is stable across CUDA-graph replays."""
    idx = _dev_index(device)
    side = _SIDE.get(idx)
    on_side = side is not None and torch.cuda.current_stream(torch.device('cuda', idx)).cuda_stream == side.cuda_stream
    t = _TICKETS.get((idx, on_side))
    if t is None:
        t = torch.zeros(4, dtype=torch.int32, device=torch.device('cuda', idx))
        _TICKETS[(idx, on_side)] = t
    return t.data_ptr()


# TPGSR_BN_FUSED=1: reduction + finalize in one launch ("last block finalizes").  Bit-identical results, MEASURED SLOWER on B200 (a graph
# node costs 0.9 us; the fence + ticket + one-block finalize cost more: config 3 10.23 vs 9.61 ms, profiles/r02/call8_stdout.log) -> off.
BN_FUSED = os.environ.get('TPGSR_BN_FUSED', '0') != '0'
# TPGSR_BN_PLANES=0: BatchNorm apply passes do not emit the next conv's operand planes (a separate tpgsr_split_bf16_flat pass does)
BN_PLANES = os.environ.get('TPGSR_BN_PLANES', '1') != '0'


def bn_stats(x2d: torch.Tensor, gamma, beta, running_mean, running_var, momentum=0.1, eps=1e-5):
    M, C = x2d.shape
    mean_rstd = new((2, C), x2d)
    scale_shift = new((2, C), x2d)
    wp, wn = workspace(query('tpgsr_colreduce_workspace_bytes', M, C), x2d.device)
    if BN_FUSED:
        _call('tpgsr_bn_stats_fused', ptr(x2d), M, C, C, ptr(gamma), ptr(beta), ptr(running_mean), ptr(running_var), momentum, eps,
              ptr(mean_rstd), ptr(scale_shift), wp, wn, ticket(x2d.device), stream())
    else:
        _call('tpgsr_bn_stats', ptr(x2d), M, C, C, ptr(gamma), ptr(beta), ptr(running_mean), ptr(running_var), momentum, eps,
              ptr(mean_rstd), ptr(scale_shift), wp, wn, stream())
    return mean_rstd, scale_shift


def bn_apply_planes(x: torch.Tensor, scale_shift: torch.Tensor, act: int, residual: Optional[torch.Tensor], P: int, need_lo: bool = True):
    """x dense NHWC [B,H,W,C] -> (y dense fp32, FlatPlanes of y with padding P): BN apply + activation (+ residual) + operand planes."""
    B, H, W, C = x.shape
    y = torch.empty_like(x)
    rows = B * (H + P) * (W + P)
    hi = torch.empty((rows, C), dtype=torch.bfloat16, device=x.device)
    lo = torch.empty((rows, C), dtype=torch.bfloat16, device=x.device) if need_lo else None
    _call('tpgsr_bn_apply_planes', ptr(x), ptr(y), ptr(scale_shift), B, H, W, C, act, ptr(residual), P, ptr(hi), ptr(lo), stream())
    return y, FlatPlanes(hi, lo, B, H, W, C, P)


def bn_bwd_fused(dy: torch.Tensor, x: torch.Tensor, scale_shift, mean_rstd, act: int, train: bool, dgamma, dbeta, accumulate: bool,
                 P: Optional[int] = None, need_lo: bool = True):
    """Dense NHWC [B,H,W,C] BN backward in two launches; with ``P`` also returns dx as padded-flat operand planes -> (dx, FlatPlanes | None)."""
    B, H, W, C = x.shape
    dx = torch.empty_like(x)
    hi = lo = None
    if P is not None:
        rows = B * (H + P) * (W + P)
        hi = torch.empty((rows, C), dtype=torch.bfloat16, device=x.device)
        lo = torch.empty((rows, C), dtype=torch.bfloat16, device=x.device) if need_lo else None
    wp, wn = workspace(query('tpgsr_bn_bwd_workspace_bytes', B * H * W, C), x.device)
    _call('tpgsr_bn_bwd_fused', ptr(dy), ptr(x), ptr(dx), ptr(scale_shift), ptr(mean_rstd), B, H, W, C, act, int(train), ptr(dgamma), ptr(dbeta),
          int(accumulate), P or 0, ptr(hi), ptr(lo), wp, wn, ticket(x.device) if BN_FUSED else None, stream())
    return dx, (FlatPlanes(hi, lo, B, H, W, C, P) if P is not None else None)


def bn_eval_prepare(gamma, beta, running_mean, running_var, eps=1e-5):
    C = gamma.numel()
    mean_rstd = new((2, C), gamma)
    scale_shift = new((2, C), gamma)
    _call('tpgsr_bn_eval_prepare', C, ptr(gamma), ptr(beta), ptr(running_mean), ptr(running_var), eps, ptr(mean_rstd),
          ptr(scale_shift), stream())
    return mean_rstd, scale_shift


def bn_apply(x2d: torch.Tensor, scale_shift: torch.Tensor, act: int = ACT_NONE, residual: Optional[torch.Tensor] = None,
             out: Optional[torch.Tensor] = None, ldy: int = 0) -> torch.Tensor:
    M, C = x2d.shape
    if out is None:
        out = new((M, C), x2d)
    _call('tpgsr_bn_apply', ptr(x2d), ptr(out), ptr(scale_shift), M, C, C, ldy or C, act, ptr(residual), C, stream())
    return out


def bn_bwd(dy2d: torch.Tensor, x2d: torch.Tensor, scale_shift, mean_rstd, act: int, train: bool, dgamma, dbeta,
           accumulate: bool, lddy: int = 0) -> torch.Tensor:
    M, C = x2d.shape
    dx = new((M, C), x2d)
    wp, wn = workspace(query('tpgsr_bn_bwd_workspace_bytes', M, C), x2d.device)
    _call('tpgsr_bn_bwd', ptr(dy2d), ptr(x2d), ptr(dx), ptr(scale_shift), ptr(mean_rstd), M, C, lddy or C, C, C, act,
          int(train), ptr(dgamma), ptr(dbeta), int(accumulate), wp, wn, stream())
    return dx


# ------------------------------------------------------------------------------------------------ activations etc.
def act_fwd(x: torch.Tensor, act: int, slope: Optional[torch.Tensor] = None) -> torch.Tensor:
    y = torch.empty_like(x)
    _call('tpgsr_act_fwd', ptr(x), ptr(y), x.numel(), act, ptr(slope), stream())
    return y


def act_bwd(dy: torch.Tensor, ref: torch.Tensor, act: int, slope: Optional[torch.Tensor] = None,
            dslope: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    dx = torch.empty_like(dy)
    wp, wn = workspace(query('tpgsr_act_bwd_workspace_bytes', dy.numel()), dy.device)
    _call('tpgsr_act_bwd', ptr(dy), ptr(ref), ptr(dx), dy.numel(), act, ptr(slope), ptr(dslope), int(accumulate), wp, wn, stream())
    return dx


def add(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    if out is None:
        out = torch.empty_like(a)
    _call('tpgsr_add', ptr(a), ptr(b), ptr(out), a.numel(), stream())
    return out


def copy_cols(src: torch.Tensor, lds: int, dst: torch.Tensor, ldd: int, M: int, C: int, alpha: float = 1.0,
              accumulate: bool = False, src_off: int = 0, dst_off: int = 0) -> None:
    """dst[:, dst_off:dst_off+C] (+)= alpha * src[:, src_off:src_off+C] on pitched [M, ld] matrices."""
    _call('tpgsr_copy_cols', src.data_ptr() + 4 * src_off, lds, dst.data_ptr() + 4 * dst_off, ldd, M, C, alpha,
          int(accumulate), stream())


def fuse_linear(w2: torch.Tensor, b2, w1: torch.Tensor, b1):
    """(w2 @ w1, w2 @ b1 + b2) for [N2,N1] x [N1,K1] weight matrices (conv 1x1 weights may be passed as [N,K,1,1])."""
    N2, N1 = w2.shape[0], w2.shape[1]
    K1 = w1.numel() // N1
    wout, bout = new((N2, K1), w2), new((N2,), w2)
    _call('tpgsr_fuse_linear', ptr(w2), ptr(b2), ptr(w1), ptr(b1), ptr(wout), ptr(bout), N2, N1, K1, stream())
    return wout, bout


def add_i64(x: torch.Tensor, v: int = 1) -> None:
    _call('tpgsr_add_i64', ptr(x), x.numel(), v, stream())


def maxpool_fwd(x: torch.Tensor, k, s, p=(0, 0), need_argmax: bool = True):
    B, Hi, Wi, C = x.shape
    Ho, Wo = (Hi + 2 * p[0] - k[0]) // s[0] + 1, (Wi + 2 * p[1] - k[1]) // s[1] + 1
    y = new((B, Ho, Wo, C), x)
    arg = new((B, Ho, Wo, C), x, torch.int32) if need_argmax else None
    _call('tpgsr_maxpool_fwd', ptr(x), ptr(y), ptr(arg), B, Hi, Wi, C, k[0], k[1], s[0], s[1], p[0], p[1], stream())
    return y, arg


def maxpool_bwd(dy: torch.Tensor, arg: torch.Tensor, in_shape, k, s, p=(0, 0)) -> torch.Tensor:
    B, Hi, Wi, C = in_shape
    dx = new(in_shape, dy)
    _call('tpgsr_maxpool_bwd', ptr(dy), ptr(arg), ptr(dx), B, Hi, Wi, C, k[0], k[1], s[0], s[1], p[0], p[1], stream())
    return dx


def pixelshuffle_mish_fwd(x: torch.Tensor) -> torch.Tensor:
    B, H, W, C4 = x.shape
    y = new((B, 2 * H, 2 * W, C4 // 4), x)
    _call('tpgsr_pixelshuffle_mish_fwd', ptr(x), ptr(y), B, H, W, C4 // 4, stream())
    return y


def pixelshuffle_mish_bwd(dy: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    B, H, W, C4 = x.shape
    dx = torch.empty_like(x)
    _call('tpgsr_pixelshuffle_mish_bwd', ptr(dy), ptr(x), ptr(dx), B, H, W, C4 // 4, stream())
    return dx


# ------------------------------------------------------------------------------------------------ GRU
# developer A/B switch: TPGSR_GRU_MODE=0 forces the warp-per-sequence FFMA recurrence kernels, 2 the tensor-core kernels (default 1: per shape)
GRU_MODE = int(os.environ.get('TPGSR_GRU_MODE', '1'))
if GRU_MODE != 1:
    LIB.tpgsr_gru_set_mode(GRU_MODE)
def _seq_geom(B: int, H: int, W: int, vertical: bool):
    """(S, T, inner, outer_stride, inner_stride, t_stride) in pixel-row units."""
    if vertical:      # sequences run down the columns (gru1 in the reference, on the transposed map)
        return B * W, H, W, H * W, 1, W
    return B * H, W, 1, W, 0, 1


def gru_fwd(gx: torch.Tensor, whh: torch.Tensor, bhh: torch.Tensor, B: int, H: int, W: int, vertical: bool,
            residual: Optional[torch.Tensor] = None, need_stash: bool = True):
    """gx [M,192] -> y [M,64] (+ ysum = y + residual), stash [M,2,4,32]."""
    M = gx.shape[0]
    y = new((M, 64), gx)
    ysum = new((M, 64), gx) if residual is not None else None
    stash = new((M, 2, 4, 32), gx) if need_stash else None
    S, T, inner, so, si, st = _seq_geom(B, H, W, vertical)
    _call('tpgsr_gru_fwd', ptr(gx), ptr(whh), ptr(bhh), ptr(y), 64, ptr(residual), 64, ptr(ysum), 64, ptr(stash), S, T,
          inner, so, si, st, stream())
    return y, ysum, stash


def gru_bwd(dy: torch.Tensor, y: torch.Tensor, stash: torch.Tensor, whh: torch.Tensor, dwhh, dbhh, dbih, accumulate: bool,
            B: int, H: int, W: int, vertical: bool) -> torch.Tensor:
    """BPTT + dW_hh + bias gradients in one call (three kernels on the current stream)."""
    M = y.shape[0]
    dgx = new((M, 192), y)
    S, T, inner, so, si, st = _seq_geom(B, H, W, vertical)
    wp, wn = workspace(query('tpgsr_gru_bwd_workspace_bytes', S), y.device)
    _call('tpgsr_gru_bwd', ptr(dy), 64, ptr(y), 64, ptr(stash), ptr(whh), ptr(dgx), ptr(dwhh), ptr(dbhh), ptr(dbih),
          int(accumulate), S, T, inner, so, si, st, wp, wn, stream())
    return dgx


def gru_bwd_recurrence(dy: torch.Tensor, y: torch.Tensor, stash: torch.Tensor, whh: torch.Tensor, B: int, H: int, W: int,
                       vertical: bool):
    """The serial BPTT only: returns (dgx [M,192], part) — ``part`` holds the per-block bias partials and is completed and reduced
    by :func:`gru_bwd_weights` (which may run on the side stream)."""
    M = y.shape[0]
    dgx = new((M, 192), y)
    S, T, inner, so, si, st = _seq_geom(B, H, W, vertical)
    part = torch.empty(query('tpgsr_gru_bwd_workspace_bytes', S), dtype=torch.uint8, device=y.device)
    _call('tpgsr_gru_bwd_recurrence', ptr(dy), 64, ptr(y), 64, ptr(stash), ptr(whh), ptr(dgx), S, T, inner, so, si, st, ptr(part),
          part.numel(), stream())
    return dgx, part


def gru_bwd_weights(dgx: torch.Tensor, y: torch.Tensor, stash: torch.Tensor, part: torch.Tensor, dwhh, dbhh, dbih, accumulate: bool,
                    B: int, H: int, W: int, vertical: bool) -> None:
    S, T, inner, so, si, st = _seq_geom(B, H, W, vertical)
    _call('tpgsr_gru_bwd_weights', ptr(dgx), ptr(y), 64, ptr(stash), ptr(dwhh), ptr(dbhh), ptr(dbih), int(accumulate), S, T, inner, so, si,
          st, ptr(part), part.numel(), stream())


# ------------------------------------------------------------------------------------------------ TPS
def tps_sample_fwd(img: torch.Tensor, ctrl: torch.Tensor, inv_kernel: torch.Tensor, reprT: torch.Tensor):
    B, H, W, C = img.shape
    out = torch.empty_like(img)
    src = new((B, H * W, 2), img)
    _call('tpgsr_tps_sample_fwd', ptr(img), ptr(ctrl), ptr(inv_kernel), ptr(reprT), ptr(out), ptr(src), B, H, W, C, stream())
    return out, src


def tps_sample_bwd(dout: torch.Tensor, img: torch.Tensor, src: torch.Tensor, inv_kernel, reprT) -> torch.Tensor:
    B, H, W, C = img.shape
    dctrl = new((B, 20, 2), img)
    _call('tpgsr_tps_sample_bwd', ptr(dout), ptr(img), ptr(src), ptr(inv_kernel), ptr(reprT), ptr(dctrl), B, H, W, C, 0, stream())
    return dctrl


# ------------------------------------------------------------------------------------------------ losses
def image_loss_fwd(out: torch.Tensor, target: torch.Tensor, gradient: bool, w0: float, w1: float, out_scale: float = 1.0,
                   loss: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    B, C, H, W = out.shape
    if loss is None:
        loss = new((1,), out)
        accumulate = False
    wp, wn = workspace(query('tpgsr_loss_workspace_bytes', out.numel()), out.device)
    _call('tpgsr_image_loss_fwd', ptr(out), ptr(target), B, C, H, W, int(gradient), w0, w1, out_scale, ptr(loss),
          int(accumulate), wp, wn, stream())
    return loss


def image_loss_bwd(out: torch.Tensor, target: torch.Tensor, gradient: bool, w0: float, w1: float, scale: float = 1.0,
                   upstream: Optional[torch.Tensor] = None, dout: Optional[torch.Tensor] = None,
                   accumulate: bool = False) -> torch.Tensor:
    B, C, H, W = out.shape
    if dout is None:
        dout = torch.empty_like(out)
        accumulate = False
    _call('tpgsr_image_loss_bwd', ptr(out), ptr(target), ptr(upstream), scale, B, C, H, W, int(gradient), w0, w1, ptr(dout),
          int(accumulate), stream())
    return dout


def semantic_loss_fwd(pred: torch.Tensor, gt: torch.Tensor, out_scale: float = 1.0, loss: Optional[torch.Tensor] = None,
                      accumulate: bool = False) -> torch.Tensor:
    if loss is None:
        loss = new((1,), pred)
        accumulate = False
    wp, wn = workspace(query('tpgsr_loss_workspace_bytes', pred.numel()), pred.device)
    _call('tpgsr_semantic_loss_fwd', ptr(pred), ptr(gt), pred.numel(), out_scale, ptr(loss), int(accumulate), wp, wn, stream())
    return loss


def semantic_loss_bwd(pred: torch.Tensor, gt: torch.Tensor, scale: float = 1.0, upstream: Optional[torch.Tensor] = None,
                      dpred: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    if dpred is None:
        dpred = torch.empty_like(pred)
        accumulate = False
    _call('tpgsr_semantic_loss_bwd', ptr(pred), ptr(gt), pred.numel(), ptr(upstream), scale, ptr(dpred), int(accumulate), stream())
    return dpred
