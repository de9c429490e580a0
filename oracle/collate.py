"""Oracle for the label one-hot collate (TEST INFRASTRUCTURE — see ``oracle/__init__.py``): numpy restatement of the label loop of
``alignCollate_realWTLAMask.__call__`` (dataset/dataset.py:1263-1323).  Pinned against the arrays the reference's own class produced
(``tests/golden/label_collate.npz``, recorded by ``oracle/gen_golden_collate.py``)."""
from __future__ import annotations

from typing import Sequence

import numpy as np

D2A = '-0123456789abcdefghijklmnopqrstuvwxyz'      # dataset/dataset.py:917-918


def label_collate(label_strs: Sequence[str]):
    """-> (label_rebatches [B, 37, 1, max_len] float32, weighted_masks int64, weighted_tics int64)."""
    a2d = {ch: i for i, ch in enumerate(D2A)}
    mats, masks, tics, max_len = [], [], [], 0
    for word in label_strs:
        word = word.lower()
        if not (len(word) <= 1 or 1 < len(word) < 15):          # dataset/dataset.py:1263-1282
            word = word[:15]
        idx = [a2d[ch] for ch in word if ch in a2d]
        if idx:
            masks.extend(idx)
            m = np.zeros((len(idx), len(D2A)), np.float32)
            m[np.arange(len(idx)), idx] = 1.0
            tics.append(1)
        else:
            masks.append(0)
            m = np.zeros((1, len(D2A)), np.float32)
            m[0, 0] = 1.0
            tics.append(0)
        mats.append(m)
        max_len = max(max_len, len(word))
    out = np.zeros((len(label_strs), max_len, len(D2A)), np.float32)
    for i, m in enumerate(mats):
        out[i, :m.shape[0]] = m
    return out[:, None].transpose(0, 3, 1, 2).copy(), np.asarray(masks, np.int64), np.asarray(tics, np.int64)
