"""Per-image input transform drop-in (reference: ``dataset/dataset.py:615-632`` ``resizeNormalize``) on one CUDA kernel.

Same constructor and call surface as the reference class — ``resizeNormalize((W, H), mask)`` is what the collate functions build
(``dataset/dataset.py:929-930,1144-1145``) and call once per image (``:1245-1249``) — plus ``batch()``, which is what a collate should call:
all crops of a batch are staged as raw RGB bytes in ONE pinned buffer, copied to the device once, and resized / normalised / masked by one
launch (``tpgsr_resize_normalize``), bit-identical to Pillow's bicubic resampler + ``ToTensor``.  Results live on the GPU (there is no CPU path)."""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

from ..ops import _call, ptr, query, stream, workspace

BICUBIC = 3          # PIL.Image.BICUBIC


def _as_rgb_bytes(img) -> np.ndarray:
    a = np.asarray(img)                      # PIL images, numpy arrays and CPU tensors all speak the array protocol
    if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3 or a.shape[0] < 1 or a.shape[1] < 1:
        raise ValueError(f'resizeNormalize expects RGB uint8 images [h, w, 3], got {a.dtype} {a.shape}')
    return np.ascontiguousarray(a)


class resizeNormalize(object):
    def __init__(self, size, mask: bool = False, interpolation=BICUBIC):
        if int(interpolation) != BICUBIC:
            raise NotImplementedError('tpgsr_b200 resizeNormalize implements the reference default: Image.BICUBIC')
        self.size = (int(size[0]), int(size[1]))        # (W, H), PIL order
        self.interpolation = interpolation
        self.mask = bool(mask)

    def batch(self, imgs: Sequence, device=None) -> torch.Tensor:
        """Sequence of RGB uint8 images (ragged sizes) -> ``[B, 3 (+1), H, W]`` float32 on ``device`` (default: the current CUDA device)."""
        if not torch.cuda.is_available():
            raise RuntimeError('tpgsr_b200 runs on CUDA only — there is no CPU fallback')
        dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        arrs = [_as_rgb_bytes(i) for i in imgs]
        B = len(arrs)
        W, H = self.size
        C = 4 if self.mask else 3
        out = torch.empty((B, C, H, W), dtype=torch.float32, device=dev)
        if B == 0:
            return out
        hs = np.array([a.shape[0] for a in arrs], dtype=np.int32)
        ws = np.array([a.shape[1] for a in arrs], dtype=np.int32)
        sizes = hs.astype(np.int64) * ws.astype(np.int64) * 3
        head = 16 * B                                           # [offsets int64 x B | hs int32 x B | ws int32 x B | pixels ...]
        offsets = head + np.concatenate([[0], np.cumsum(sizes)[:-1]])
        total = head + int(sizes.sum())
        host = torch.empty(total, dtype=torch.uint8, pin_memory=True)
        hv = host.numpy()
        hv[:8 * B] = offsets.astype(np.int64).view(np.uint8)
        hv[8 * B:12 * B] = hs.view(np.uint8)
        hv[12 * B:16 * B] = ws.view(np.uint8)
        for a, o in zip(arrs, offsets):
            hv[o:o + a.size] = a.reshape(-1)
        with torch.cuda.device(dev):
            buf = host.to(dev, non_blocking=True)               # the only host->device transfer: 3 bytes per source pixel + 16 per image
            max_h, max_w = int(hs.max()), int(ws.max())
            wp, wn = workspace(query('tpgsr_resize_normalize_workspace_bytes', B, max_h, W), dev)
            base = buf.data_ptr()
            _call('tpgsr_resize_normalize', base, base, base + 8 * B, base + 12 * B, B, max_h, max_w, H, W, int(self.mask), ptr(out), wp, wn,
                  stream())
            buf.record_stream(torch.cuda.current_stream(dev))
        return out

    def __call__(self, img) -> torch.Tensor:
        return self.batch([img])[0]
