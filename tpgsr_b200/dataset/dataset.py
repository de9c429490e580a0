"""Per-image input transform drop-in (reference: ``dataset/dataset.py:615-632`` ``resizeNormalize``) on one CUDA kernel.

Same constructor and call surface as the reference class — ``resizeNormalize((W, H), mask)`` is what the collate functions build
(``dataset/dataset.py:929-930,1144-1145``) and call once per image (``:1245-1249``) — plus ``batch()``, which is what a collate should call:
all crops of a batch are staged as raw RGB bytes in ONE pinned buffer, copied to the device once, and resized / normalised / masked by one
launch (``tpgsr_resize_normalize``), bit-identical to Pillow's bicubic resampler + ``ToTensor``.  Results live on the GPU (there is no CPU path).

``stage()`` (host only) and ``launch()`` (device only) are the two halves of ``batch()``; ``tpgsr_b200.dataset.loader.BatchPrefetcher`` runs
the first on a worker thread and the second on a copy stream, one batch ahead of the train step."""
from __future__ import annotations

from typing import NamedTuple, Sequence

import numpy as np
import torch

from ..ops import _call, ptr, query, stream

BICUBIC = 3          # PIL.Image.BICUBIC


def _as_rgb_bytes(img) -> np.ndarray:
    a = np.asarray(img)                      # PIL images, numpy arrays and CPU tensors all speak the array protocol
    if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3 or a.shape[0] < 1 or a.shape[1] < 1:
        raise ValueError(f'resizeNormalize expects RGB uint8 images [h, w, 3], got {a.dtype} {a.shape}')
    return np.ascontiguousarray(a)


class Staged(NamedTuple):
    """One batch of ragged crops as the kernel reads it: ``host`` = [offsets int64 x B | heights int32 x B | widths int32 x B | pixels ...]."""
    host: torch.Tensor       # uint8, pinned when CUDA is available
    B: int
    max_h: int
    max_w: int


def stage(imgs: Sequence) -> Staged:
    """Host half: pack the crops' bytes and the per-crop table into one (pinned) buffer.  No device work."""
    arrs = [_as_rgb_bytes(i) for i in imgs]
    B = len(arrs)
    if B == 0:
        return Staged(torch.empty(0, dtype=torch.uint8), 0, 0, 0)
    hs = np.array([a.shape[0] for a in arrs], dtype=np.int32)
    ws = np.array([a.shape[1] for a in arrs], dtype=np.int32)
    sizes = hs.astype(np.int64) * ws.astype(np.int64) * 3
    head = 16 * B
    offsets = (head + np.concatenate([[0], np.cumsum(sizes)[:-1]])).astype(np.int64)      # relative to the start of the buffer
    total = head + int(sizes.sum())
    host = torch.empty(total, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
    hv = host.numpy()
    hv[:8 * B] = offsets.view(np.uint8)
    hv[8 * B:12 * B] = hs.view(np.uint8)
    hv[12 * B:16 * B] = ws.view(np.uint8)
    for a, o in zip(arrs, offsets):
        hv[o:o + a.size] = a.reshape(-1)
    return Staged(host, B, int(hs.max()), int(ws.max()))


class resizeNormalize(object):
    def __init__(self, size, mask: bool = False, interpolation=BICUBIC):
        if int(interpolation) != BICUBIC:
            raise NotImplementedError('tpgsr_b200 resizeNormalize implements the reference default: Image.BICUBIC')
        self.size = (int(size[0]), int(size[1]))        # (W, H), PIL order
        self.interpolation = interpolation
        self.mask = bool(mask)

    def launch(self, st: Staged, device=None) -> torch.Tensor:
        """Device half: ONE host->device copy (3 bytes per source pixel + 16 per crop) and ONE kernel on the current stream."""
        if not torch.cuda.is_available():
            raise RuntimeError('tpgsr_b200 runs on CUDA only — there is no CPU fallback')
        dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        W, H = self.size
        B = st.B
        out = torch.empty((B, 4 if self.mask else 3, H, W), dtype=torch.float32, device=dev)
        if B == 0:
            return out
        with torch.cuda.device(dev):
            buf = st.host.to(dev, non_blocking=True)
            # private scratch (not ops.workspace): this may run on a copy stream next to a train step that owns the shared one
            nbytes = query('tpgsr_resize_normalize_workspace_bytes', B, st.max_h, W)
            tmp = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev)
            base = buf.data_ptr()
            _call('tpgsr_resize_normalize', base, base, base + 8 * B, base + 12 * B, B, st.max_h, st.max_w, H, W, int(self.mask), ptr(out),
                  ptr(tmp), nbytes, stream())
        return out

    def batch(self, imgs: Sequence, device=None) -> torch.Tensor:
        """Sequence of RGB uint8 images (ragged sizes) -> ``[B, 3 (+1), H, W]`` float32 on ``device`` (default: the current CUDA device)."""
        if not torch.cuda.is_available():
            raise RuntimeError('tpgsr_b200 runs on CUDA only — there is no CPU fallback')
        return self.launch(stage(imgs), device)

    def __call__(self, img) -> torch.Tensor:
        return self.batch([img])[0]
