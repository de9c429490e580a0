"""ORACLE (test infrastructure, never imported by the product path): CPU restatement of the reference's per-image input transform.

Follows ``dataset/dataset.py:615-632`` (``resizeNormalize.__call__``): ``img.resize(size, Image.BICUBIC)`` -> ``ToTensor`` -> optional mask
channel ``img.convert('L')``, threshold at the mean, ``0 if x > thres else 255`` -> ``ToTensor`` -> ``torch.cat``.

The arithmetic lives in a third-party dependency that is absent from ``/root/reference``: **Pillow** (no version pinned by the reference;
12.2.0 in this image — the 8-bit resampler has been unchanged since Pillow 3.x/7.x) and torchvision's ``ToTensor``.  Restated from Pillow's
published algorithm (``src/libImaging/Resample.c``: ``precompute_coeffs``, ``normalize_coeffs_8bpc``, ``ImagingResampleHorizontal_8bpc`` /
``Vertical_8bpc``; ``src/libImaging/Convert.c``: ``L24``):

* per output coordinate ``xx``: ``scale = in/out``, ``filterscale = max(scale, 1)``, ``support = 2 * filterscale`` (bicubic, a = -0.5),
  ``center = (xx + 0.5) * scale``, ``xmin = max(int(center - support + 0.5), 0)``, ``xmax = min(int(center + support + 0.5), in)``, weights
  ``w[x] = bicubic((x + xmin - center + 0.5) / filterscale)`` normalised by their sum (all in float64), then rounded to 22-bit fixed point;
* horizontal pass first, then vertical, each ``clip8((2^21 + sum(pixel * k)) >> 22)`` with a uint8 image in between; a pass whose
  input and output size agree is skipped;
* ``L = (19595 R + 38470 G + 7471 B + 0x8000) >> 16``; the mask compares integers with the float64 mean;
* ``ToTensor``: uint8 -> float32 / 255, CHW.

Pinned here against Pillow itself (``oracle/gen_golden_resize.py`` records ``tests/golden/resize_normalize.npz`` by running the reference's own
class on seeded images; ``tests/test_oracle_golden.py`` additionally compares with the live Pillow when it is importable).  Bit-exact.
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def bicubic(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def coeffs(in_size: int, out_size: int):
    """(ksize, bounds [out,2] int32 (xmin, count), kk [out,ksize] int32 fixed-point weights) — Resample.c precompute + normalize."""
    scale = float(np.float32(in_size) - np.float32(0)) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)            # C cast: truncation toward zero
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + k * (1 << PRECISION_BITS)) if k < 0 else int(0.5 + k * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _pass(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    """One resampling pass of a uint8 [h,w,c] image along ``axis`` (1 = horizontal, 0 = vertical)."""
    in_size = img.shape[axis]
    if in_size == out_size:
        return img
    _, bounds, kk = coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)                   # [in, other, c]
    out = np.empty((out_size,) + src.shape[1:], dtype=np.uint8)
    for xx in range(out_size):
        xmin, n = bounds[xx]
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(kk[xx, :n].astype(np.int64), src[xmin:xmin + n], axes=(0, 0))
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bicubic_u8(img: np.ndarray, size: Tuple[int, int]) -> np.ndarray:
    """PIL ``Image.resize(size=(W, H), Image.BICUBIC)`` of a uint8 RGB image [h,w,3]."""
    W, H = size
    return np.ascontiguousarray(_pass(_pass(img, W, 1), H, 0))


def luma_u8(rgb: np.ndarray) -> np.ndarray:
    r, g, b = (rgb[..., i].astype(np.int64) for i in range(3))
    return ((r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16).astype(np.uint8)


def resize_normalize(img: np.ndarray, size: Tuple[int, int], mask: bool) -> np.ndarray:
    """uint8 RGB [h,w,3] -> float32 [3 (+1), H, W] in [0,1]; channel 3 = 1.0 where the luma is <= its mean, else 0.0."""
    out = resize_bicubic_u8(img, size)
    chans = [out[..., c] for c in range(3)]
    if mask:
        L = luma_u8(out)
        thres = np.array(L).mean()
        chans.append(np.where(L > thres, 0, 255).astype(np.uint8))
    return (np.stack(chans).astype(np.float32) / np.float32(255)).astype(np.float32)
