"""Label half of the reference's collate functions (``dataset/dataset.py:1239-1323`` ``alignCollate_realWTLAMask.__call__``; the image half is
``tpgsr_b200.dataset.resizeNormalize.batch``): label strings -> ``label_rebatches [B, 37, 1, max_len]`` one-hot over
``"-0123456789abcdefghijklmnopqrstuvwxyz"``, ``weighted_masks`` (the class indices of all labels back to back; one 0 for an empty label) and
``weighted_tics`` (1 where a label has at least one in-alphabet character).  Characters are mapped on the host, as in the reference;
the batch tensor is built by one kernel from one int32 array (``tpgsr_label_onehot``) instead of a Python loop of scatter_ calls."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

from ..ops import _call, ptr, stream

ALPHABET = '-0123456789abcdefghijklmnopqrstuvwxyz'          # d2a of alignCollate_syn.__init__ (dataset/dataset.py:917-923)
_A2D = {ch: i for i, ch in enumerate(ALPHABET)}


def encode_labels(label_strs: Sequence[str], max_chars: int = 15) -> Tuple[np.ndarray, List[int], List[int], int]:
    """Host half: (codes [B, max_len] int32 with -1 padding, weighted_masks, weighted_tics, max_len).  ``max_chars``: words of 15 characters or
    more are cut to 15 (dataset/dataset.py:1263-1282); characters outside the alphabet are skipped; ``max_len`` is the longest (cut) WORD, as
    in the reference — not the longest code list."""
    lists, masks, tics, max_len = [], [], [], 0
    for word in label_strs:
        word = word.lower()
        if len(word) >= max_chars:
            word = word[:max_chars]
        codes = [_A2D[ch] for ch in word if ch in _A2D]
        if codes:
            masks.extend(codes)
            tics.append(1)
        else:
            masks.append(0)
            tics.append(0)
            codes = [0]                                   # "assign a blank label" (dataset/dataset.py:1305-1308)
        lists.append(codes)
        max_len = max(max_len, len(word))
    if any(len(c) > max_len for c in lists):
        raise ValueError('label collate: a batch of empty labels has max_len 0 (the reference fails on the same batch, dataset/dataset.py:1313-1314)')
    arr = np.full((len(lists), max(max_len, 1)), -1, dtype=np.int32)
    for i, c in enumerate(lists):
        arr[i, :len(c)] = c
    return arr, masks, tics, max_len


def label_onehot_collate(label_strs: Sequence[str], device=None, max_chars: int = 15):
    """-> (label_rebatches [B, 37, 1, max_len] float32 on the GPU, weighted_masks int64 [sum], weighted_tics int64 [B])."""
    if not torch.cuda.is_available():
        raise RuntimeError('tpgsr_b200 runs on CUDA only — there is no CPU fallback')
    dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    codes, masks, tics, max_len = encode_labels(label_strs, max_chars)
    B = len(label_strs)
    out = torch.empty((B, len(ALPHABET), 1, max_len), dtype=torch.float32, device=dev)
    if B and max_len:
        with torch.cuda.device(dev):
            host = torch.from_numpy(codes)
            d_codes = (host.pin_memory() if B * max_len > 4096 else host).to(dev, non_blocking=True)
            _call('tpgsr_label_onehot', ptr(d_codes), B, max_len, len(ALPHABET), ptr(out), stream())
    return out, torch.tensor(masks, dtype=torch.long), torch.tensor(tics, dtype=torch.long)
