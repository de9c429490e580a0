"""Golden fixture for the per-image input transform: imports the reference's own ``dataset/dataset.py`` (read-only,
``/root/reference``; ``lmdb``, ``pyfasttext``, ``IPython``, ``editdistance``, ``matplotlib`` / ``mpl_toolkits`` are absent here and not used by the class, so
they are stubbed), runs ``resizeNormalize((W, H), mask)`` — Pillow bicubic resize, ToTensor, mean-threshold mask — on seeded ragged RGB
crops and stores inputs + outputs in ``tests/golden/resize_normalize.npz``.

    python -m oracle.gen_golden_resize
"""
from __future__ import annotations

import sys
import types

import numpy as np

from oracle.gen_golden import REF, save


def _import_dataset():
    for name in ('IPython', 'editdistance', 'lmdb', 'pyfasttext', 'matplotlib', 'matplotlib.pyplot', 'mpl_toolkits', 'mpl_toolkits.mplot3d'):
        mod = types.ModuleType(name)
        mod.embed = lambda *a, **k: None
        mod.FastText = object
        mod.Axes3D = object
        sys.modules.setdefault(name, mod)
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    sys.modules['mpl_toolkits'].mplot3d = sys.modules['mpl_toolkits.mplot3d']
    import scipy
    if not hasattr(scipy, 'finfo'):          # dataset/dataset.py:33 builds a blur kernel at import time through scipy.finfo (removed in SciPy 1.x)
        scipy.finfo = np.finfo
    sys.path.insert(0, REF)
    import importlib
    return importlib.import_module('dataset.dataset')


def synth_crops(seed: int, n: int):
    """Ragged uint8 RGB crops shaped like TextZoom's (a few pixels up to ~90 x 300), three textures: noise, smooth blobs, binary text-like."""
    r = np.random.RandomState(seed)
    fixed = [(1, 1), (2, 3), (16, 64), (32, 128), (16, 150), (50, 64), (3, 1), (1, 5), (75, 240), (17, 63), (32, 127), (15, 64)]
    crops = []
    for i in range(n):
        h, w = fixed[i] if i < len(fixed) else (r.randint(4, 60), r.randint(8, 200))
        kind = i % 3
        if kind == 0:
            a = r.rand(h, w, 3) * 255
        elif kind == 1:
            yy, xx = np.mgrid[0:h, 0:w]
            a = np.stack([128 + 100 * np.sin(xx / (3.0 + c) + c) * np.cos(yy / (2.0 + c)) for c in range(3)], -1) + r.randn(h, w, 3) * 8
        else:
            a = (r.rand(h, w, 1) > 0.6) * np.array([[[230.0, 220.0, 240.0]]]) + 10
        crops.append(np.clip(a, 0, 255).astype(np.uint8))
    return crops


def main():
    from PIL import Image
    ds = _import_dataset()
    crops = synth_crops(77, 20)
    out = {'n': np.array(len(crops))}
    for (W, H, tag) in ((128, 32, 'hr'), (64, 16, 'lr')):
        for mask in (False, True):
            tf = ds.resizeNormalize((W, H), mask)
            res = np.stack([tf(Image.fromarray(c, 'RGB')).numpy() for c in crops])
            # ToTensor is uint8 / 255 in float32: store the bytes (4x smaller) after checking that they reproduce the floats exactly
            u8 = np.rint(res * 255).astype(np.uint8)
            assert res.dtype == np.float32 and np.array_equal(u8.astype(np.float32) / np.float32(255), res)
            out[f'y/{tag}/{int(mask)}'] = u8
    for i, c in enumerate(crops):
        out[f'x/{i}'] = c
    import PIL
    out['pillow_version'] = np.array(PIL.__version__)
    save('resize_normalize', **out)


if __name__ == '__main__':
    main()
