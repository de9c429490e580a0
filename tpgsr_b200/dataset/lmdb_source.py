"""TextZoom LMDB sample source (reference: ``dataset/dataset.py:104-149`` ``lmdbDataset_real``; ``buf2PIL`` ``:43-50``): the key / value layout
of the TextZoom databases — ``num-samples``, ``label-%09d``, ``image_hr-%09d``, ``image_lr-%09d`` (1-based), images as encoded files — read
into ``(image_hr, image_lr, label_str)`` samples, the iterable ``tpgsr_b200.dataset.BatchPrefetcher`` consumes.

Host-side I/O only (as in the reference).  ``lmdb`` is a third-party package that this image does not ship: a database PATH needs it at run
time; any object with the ``env.begin(write=False) -> txn.get(key)`` protocol works as well (``DictEnv`` below is the in-memory one the
tests use), so the reader itself carries no dependency."""
from __future__ import annotations

import io
from typing import Dict, Iterator, Tuple

from ..utils.util import str_filt


class DictEnv:
    """Minimal stand-in for an ``lmdb.Environment`` over a dict of ``bytes -> bytes``."""

    def __init__(self, data: Dict[bytes, bytes]):
        self.data = dict(data)

    class _Txn:
        def __init__(self, data):
            self._d = data

        def get(self, key):
            return self._d.get(key)

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False

    def begin(self, write: bool = False):
        return DictEnv._Txn(self.data)


def _open(root):
    if hasattr(root, 'begin'):
        return root
    try:
        import lmdb
    except ImportError as e:
        raise ImportError('reading an LMDB path needs the third-party `lmdb` package (not part of this image); pass an environment object '
                          'with the env.begin().get(key) protocol instead') from e
    env = lmdb.open(root, max_readers=1, readonly=True, lock=False, readahead=False, meminit=False)      # dataset/dataset.py:107-113
    if not env:
        raise IOError('cannot open lmdb at %s' % root)
    return env


def buf2PIL(txn, key: bytes, mode: str = 'RGB'):
    from PIL import Image
    buf = txn.get(key)
    if buf is None:
        raise KeyError(key)
    return Image.open(io.BytesIO(buf)).convert(mode)


class lmdbDataset_real:
    """``lmdbDataset_real(root, voc_type, max_len, test)``: ``len()``, ``[index] -> (img_HR, img_lr, label_str)``, and iteration."""

    def __init__(self, root=None, voc_type: str = 'upper', max_len: int = 100, test: bool = False):
        self.env = _open(root)
        with self.env.begin(write=False) as txn:
            self.nSamples = int(txn.get(b'num-samples'))
        self.voc_type, self.max_len, self.test = voc_type, max_len, test

    def __len__(self) -> int:
        return self.nSamples

    def __getitem__(self, index: int) -> Tuple[object, object, str]:
        if not 0 <= index < self.nSamples:
            raise IndexError('index range error')
        for probe in range(self.nSamples):                 # an undecodable sample is replaced by the next one, as in the reference
            i = (index + probe) % self.nSamples + 1
            txn = self.env.begin(write=False)
            word = str(txn.get(b'label-%09d' % i).decode())
            try:
                img_hr = buf2PIL(txn, b'image_hr-%09d' % i, 'RGB')
                img_lr = buf2PIL(txn, b'image_lr-%09d' % i, 'RGB')
            except (IOError, KeyError):
                continue
            return img_hr, img_lr, str_filt(word, self.voc_type)
        raise IOError('no decodable sample in the database')

    def __iter__(self) -> Iterator[Tuple[object, object, str]]:
        for i in range(self.nSamples):
            yield self[i]
