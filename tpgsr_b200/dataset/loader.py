"""Batches on the GPU one step ahead of the train step.

The reference feeds ``TextSR.train`` from a ``torch.utils.data.DataLoader`` whose collate (``dataset/dataset.py:1239-1260``) runs Pillow's
``resizeNormalize`` per image and ``torch.cat``s fp32 tensors, which the loop then moves with ``.to(self.device)``
(``interfaces/super_resolution.py:306-308``).  Here a worker thread only PACKS raw crop bytes into pinned buffers (``dataset.stage``); the
consumer enqueues, on a private copy stream, one host->device copy and one resize kernel per resolution (``resizeNormalize.launch``) for batch
k + 1 before it hands out batch k, and the compute stream waits on an event — so decode/packing, PCIe and the resize overlap the step.

``samples`` is any iterable of ``(image_hr, image_lr, label_str)`` with RGB uint8 images (PIL / numpy); the LMDB reader of the reference
(``dataset/dataset.py:104-149``) is such an iterable, but ``lmdb`` is not part of this image, so the source stays pluggable."""
from __future__ import annotations

import queue
import threading
from typing import Iterable, Iterator, List, Optional, Tuple

import torch

from .dataset import Staged, resizeNormalize, stage

_END = object()


def batches(samples: Iterable, batch_size: int, drop_last: bool = True) -> Iterator[Tuple[list, list, list]]:
    """Group ``(hr, lr, label)`` samples into lists of ``batch_size`` (host only)."""
    hr, lr, lab = [], [], []
    for s in samples:
        hr.append(s[0]); lr.append(s[1]); lab.append(s[2])
        if len(hr) == batch_size:
            yield hr, lr, lab
            hr, lr, lab = [], [], []
    if hr and not drop_last:
        yield hr, lr, lab


class BatchPrefetcher:
    """Iterate ``(images_hr [B,C,2H,2W], images_lr [B,C,H,W], label_strs)`` with the tensors already on ``device``.

    ``size_hr`` = (W, H) of the HR tensor (reference default (128, 32), ``imgW x imgH`` of ``alignCollate_real``); the LR tensor is
    ``size_hr // down_sample_scale``.  ``depth`` staged batches are kept ahead by the worker thread; one batch is kept ahead on the device."""

    def __init__(self, samples: Iterable, batch_size: int, size_hr=(128, 32), down_sample_scale: int = 2, mask: bool = True, device=None,
                 depth: int = 2, drop_last: bool = True, _launch=None):
        self.samples, self.batch_size, self.drop_last = samples, int(batch_size), drop_last
        self.tf_hr = resizeNormalize(size_hr, mask)
        self.tf_lr = resizeNormalize((size_hr[0] // down_sample_scale, size_hr[1] // down_sample_scale), mask)
        self.device = device
        self.depth = max(1, int(depth))
        self._launch = _launch          # test hook: (transform, Staged) -> tensor, replaces the device half

    # ---- worker thread: host-only packing ---------------------------------------------------------------------------------
    def _produce(self, q: 'queue.Queue', stop: threading.Event) -> None:
        def put(item) -> bool:                # False once the consumer has gone away
            while not stop.is_set():
                try:
                    q.put(item, timeout=0.1)
                    return True
                except queue.Full:
                    continue
            return False

        try:
            for hr, lr, lab in batches(self.samples, self.batch_size, self.drop_last):
                if not put((stage(hr), stage(lr), lab)):
                    return
            put(_END)
        except BaseException as e:          # surface worker errors in the consumer instead of hanging it
            put(e)

    # ---- consumer: device half, one batch ahead -------------------------------------------------------------------------------
    def _to_device(self, item, copy_stream):
        st_hr, st_lr, lab = item
        if self._launch is not None:
            return self._launch(self.tf_hr, st_hr), self._launch(self.tf_lr, st_lr), lab, None
        with torch.cuda.stream(copy_stream):
            hr = self.tf_hr.launch(st_hr, self.device)
            lr = self.tf_lr.launch(st_lr, self.device)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return hr, lr, lab, ev

    def __iter__(self):
        q: 'queue.Queue' = queue.Queue(maxsize=self.depth)
        stop = threading.Event()
        worker = threading.Thread(target=self._produce, args=(q, stop), daemon=True)
        worker.start()
        copy_stream = None
        if self._launch is None:
            if not torch.cuda.is_available():
                raise RuntimeError('tpgsr_b200 runs on CUDA only — there is no CPU fallback')
            dev = torch.device('cuda', torch.cuda.current_device()) if self.device is None else torch.device(self.device)
            copy_stream = torch.cuda.Stream(device=dev)

        def fetch():
            item = q.get()
            if item is _END:
                return None
            if isinstance(item, BaseException):
                raise item
            return self._to_device(item, copy_stream)

        try:
            ahead = fetch()
            while ahead is not None:
                cur = ahead
                ahead = fetch()                       # batch k + 1 is enqueued on the copy stream before batch k is handed out
                hr, lr, lab, ev = cur
                if ev is not None:
                    main = torch.cuda.current_stream(hr.device)
                    main.wait_event(ev)
                    hr.record_stream(main)
                    lr.record_stream(main)
                yield hr, lr, lab
        finally:
            stop.set()
