"""Golden fixture for the label one-hot collate: imports the reference's own ``dataset/dataset.py`` (read-only,
``/root/reference``; lmdb / pyfasttext / IPython / matplotlib are absent here and only needed by other classes of that file, so they are stubbed, and the
scipy aliases its import-time code still calls are restored), runs ``alignCollate_realWTLAMask.__call__`` on seeded label strings with tiny
dummy images, and stores the strings + the three label outputs in ``tests/golden/label_collate.npz``.

    python -m oracle.gen_golden_collate
"""
from __future__ import annotations

import sys
import types

import numpy as np

from oracle.gen_golden import REF, save


def _import_dataset():
    for name, attrs in (('IPython', dict(embed=lambda *a, **k: None)), ('lmdb', {}), ('pyfasttext', dict(FastText=object)),
                        ('matplotlib', {}), ('matplotlib.pyplot', {}), ('mpl_toolkits', {}), ('mpl_toolkits.mplot3d', dict(Axes3D=object))):
        try:
            __import__(name)
            continue
        except ImportError:
            pass
        mod = types.ModuleType(name)
        mod.__dict__.update(attrs)
        mod.__path__ = []
        sys.modules[name] = mod
    import scipy
    for alias in ('finfo', 'pi', 'exp', 'sqrt'):
        if not hasattr(scipy, alias):
            setattr(scipy, alias, getattr(np, alias))
    sys.path.insert(0, REF)
    import importlib
    return importlib.import_module('dataset.dataset')


def main():
    from PIL import Image
    ds = _import_dataset()
    r = np.random.RandomState(11)
    pool = 'abcdefghijklmnopqrstuvwxyz0123456789ABCXYZ-_!? '
    words = [''.join(pool[i] for i in r.randint(0, len(pool), size=n)) for n in (1, 2, 5, 9, 14, 15, 16, 23, 3, 7)]
    words += ['', '!!', 'Hello', 'a', 'WORLD2024', 'x' * 30]
    collate = ds.alignCollate_realWTLAMask(imgH=32, imgW=128, down_sample_scale=2, mask=True)
    img_hr = Image.fromarray(r.randint(0, 255, (32, 128, 3)).astype(np.uint8))
    img_lr = Image.fromarray(r.randint(0, 255, (16, 64, 3)).astype(np.uint8))
    rec = {}
    for name, sel in (('all', list(range(len(words)))), ('short', [0, 1, 8, 13]), ('with_empty', [10, 11, 12])):
        batch = [(img_hr, img_lr, words[i]) for i in sel]
        _, _, strs, rebatches, masks, tics = collate(batch)
        rec[name + '/strs'] = np.array(list(strs))
        rec[name + '/label_rebatches'] = rebatches.numpy()
        rec[name + '/weighted_masks'] = masks.numpy()
        rec[name + '/weighted_tics'] = tics.numpy()
    save('label_collate', **rec)


if __name__ == '__main__':
    main()
