"""Golden fixture for the CTC greedy decode: imports the reference's own ``utils/metrics.py`` / ``utils/util.py``
(read-only, ``/root/reference``; ``editdistance`` and ``IPython`` are absent here and only needed by other functions of those files, so
they are stubbed), runs ``get_string_crnn`` / ``str_filt`` on seeded inputs and stores inputs + outputs in ``tests/golden/ctc_decode.npz``.

    python -m oracle.gen_golden_ctc
"""
from __future__ import annotations

import sys
import types

import numpy as np
import torch

from oracle.gen_golden import REF, save


def _import_metrics():
    for name in ('IPython', 'editdistance'):
        mod = types.ModuleType(name)
        mod.embed = lambda *a, **k: None
        sys.modules.setdefault(name, mod)
    sys.path.insert(0, REF)
    import importlib
    return importlib.import_module('utils.metrics'), importlib.import_module('utils.util')


def _peaked(r, T, B, C, p_blank=0.35, p_repeat=0.35, noise=1.0):
    """Logits whose arg-max path has blanks and repeats the way a trained CRNN's does."""
    path = np.zeros((T, B), dtype=np.int64)
    for b in range(B):
        cur = 0
        for t in range(T):
            u = r.rand()
            if u < p_blank:
                cur = 0
            elif u < p_blank + p_repeat and t > 0:
                cur = path[t - 1, b]
            else:
                cur = r.randint(1, C)
            path[t, b] = cur
    x = r.randn(T, B, C).astype(np.float32) * noise
    for t in range(T):
        for b in range(B):
            x[t, b, path[t, b]] += 6.0
    return x


def main():
    metrics, util = _import_metrics()
    r = np.random.RandomState(2024)
    cases = {}
    cases['crnn'] = _peaked(r, 26, 48, 37)                          # the CRNN's output shape at B=48 (model/crnn/crnn.py:74-90)
    cases['noise'] = r.randn(26, 7, 37).astype(np.float32)          # unpeaked: almost no repeats
    cases['long'] = _peaked(r, 70, 5, 37, 0.2, 0.6)                 # T > 32 (more than one warp chunk), long runs
    cases['t1'] = _peaked(r, 1, 9, 37, 0.5, 0.0)
    edge = np.full((8, 6, 37), -1.0, dtype=np.float32)
    edge[:, 0, 0] = 1.0                                             # all blank -> ''
    edge[:, 1, 5] = 1.0                                             # one character repeated -> '4'
    edge[0::2, 2, 11] = 1.0; edge[1::2, 2, 0] = 1.0                 # a-a-a- -> 'aaaa' (blank separates repeats)
    edge[:, 3, 36] = 1.0; edge[7, 3, 36] = -1.0; edge[7, 3, 1] = 1.0    # last class, then a change on the last step
    edge[:, 4, :] = 0.25                                            # exact ties everywhere -> index 0 -> ''
    edge[:, 5, :] = -3.0; edge[:, 5, 3] = 2.0; edge[:, 5, 9] = 2.0; edge[4:, 5, 2] = 2.0   # ties between classes: the lowest index wins
    cases['edge'] = edge
    out = {}
    for k, x in cases.items():
        strs = metrics.get_string_crnn(torch.from_numpy(x))
        out['x/' + k] = x
        out['s/' + k] = np.array(strs, dtype='U80')
    labels = ['Hello, World!', 'ABC-123', '', 'a.b_c d', 'UPPER lower 09', '---', 'é1z']
    for voc in ('digit', 'lower', 'upper', 'all'):
        out['filt/' + voc] = np.array([util.str_filt(s, voc) for s in labels], dtype='U40')
    out['labels'] = np.array(labels, dtype='U40')
    save('ctc_decode', **out)
    print({k: list(v[:4]) for k, v in out.items() if k.startswith('s/')})


if __name__ == '__main__':
    main()
