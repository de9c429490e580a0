"""ORACLE (test infrastructure, never imported by the product path): CPU restatement of the reference's CTC greedy decode.

Follows ``utils/metrics.py:71-88`` (``get_string_crnn``): per sample, ``torch.max(output, 1)[1]`` over the class axis for every time step,
then a left-to-right scan that appends ``alphabet[i]`` unless it equals the last appended character, with the blank (index 0) resetting
"last" — i.e. collapse repeats, then drop blanks.  The reference compares CHARACTERS (``alphabet[i] != last``); with an alphabet of
distinct characters (the default ``'-0123456789abcdefghijklmnopqrstuvwxyz'``, also what ``interfaces/base.py:634-636`` builds the CRNN for)
that is the same as comparing indices, which is what ``greedy_ids`` does.  Ties in the arg-max resolve to the lowest index (what the CPU
``torch.max`` the fixtures were recorded with does).  ``str_filt`` restates ``utils/util.py:12-24``.

Pinned by ``tests/golden/ctc_decode.npz`` (recorded from the reference's own function by ``oracle/gen_golden_ctc.py``).
"""
from __future__ import annotations

import string
from typing import List, Tuple

import numpy as np

ALPHABET = '-0123456789abcdefghijklmnopqrstuvwxyz'


def greedy_ids(logits: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """logits [T,B,C] -> (ids [B,T] int32 padded with -1, lens [B] int32)."""
    T, B, C = logits.shape
    ids = np.full((B, T), -1, dtype=np.int32)
    lens = np.zeros((B,), dtype=np.int32)
    arg = np.argmax(logits, axis=2)            # first maximal index, [T,B]
    for b in range(B):
        last = 0
        n = 0
        for t in range(T):
            i = int(arg[t, b])
            if i != last and i != 0:
                ids[b, n] = i
                n += 1
            last = i
        lens[b] = n
    return ids, lens


def get_string_crnn(outputs, alphabet: str = ALPHABET) -> List[str]:
    logits = np.asarray(outputs, dtype=np.float32)
    ids, lens = greedy_ids(logits)
    return [''.join(alphabet[i] for i in ids[b, :lens[b]]) for b in range(logits.shape[1])]


def str_filt(str_: str, voc_type: str) -> str:
    alpha = {'digit': string.digits, 'lower': string.digits + string.ascii_lowercase, 'upper': string.digits + string.ascii_letters,
             'all': string.digits + string.ascii_letters + string.punctuation}[voc_type]
    if voc_type == 'lower':
        str_ = str_.lower()
    return ''.join(c for c in str_ if c in alpha)
