"""Host logic of the input pipeline (no GPU): packing of ragged crops into the kernel's staging buffer, batching, and the prefetcher's
ordering / shutdown / error behaviour with the device half stubbed out."""
import threading

import numpy as np
import pytest
import torch


def _crops(n, seed=0):
    r = np.random.RandomState(seed)
    return [(r.rand(r.randint(1, 12), r.randint(1, 20), 3) * 255).astype(np.uint8) for _ in range(n)]


def test_stage_layout_round_trips():
    from tpgsr_b200.dataset import stage
    crops = _crops(7)
    st = stage(crops)
    raw = st.host.numpy()
    B = st.B
    assert B == 7 and st.max_h == max(c.shape[0] for c in crops) and st.max_w == max(c.shape[1] for c in crops)
    offsets = raw[:8 * B].view(np.int64)
    hs = raw[8 * B:12 * B].view(np.int32)
    ws = raw[12 * B:16 * B].view(np.int32)
    assert offsets[0] == 16 * B and raw.size == 16 * B + sum(c.size for c in crops)
    for i, c in enumerate(crops):
        assert (hs[i], ws[i]) == c.shape[:2]
        assert np.array_equal(raw[offsets[i]:offsets[i] + c.size].reshape(c.shape), c)
    assert stage([]).B == 0
    for bad in (np.zeros((4, 4), np.uint8), np.zeros((4, 4, 3), np.float32), np.zeros((0, 4, 3), np.uint8), np.zeros((4, 4, 4), np.uint8)):
        with pytest.raises(ValueError):
            stage([bad])
    # non-contiguous views and anything with the array protocol are accepted
    big = _crops(1, 3)[0]
    assert np.array_equal(stage([big[:, ::-1]]).host.numpy()[16:], np.ascontiguousarray(big[:, ::-1]).reshape(-1))


def test_batches_grouping():
    from tpgsr_b200.dataset import batches
    samples = [(f'h{i}', f'l{i}', str(i)) for i in range(7)]
    got = list(batches(samples, 3))
    assert [b[2] for b in got] == [['0', '1', '2'], ['3', '4', '5']]
    got = list(batches(samples, 3, drop_last=False))
    assert [b[2] for b in got][-1] == ['6'] and got[0][0] == ['h0', 'h1', 'h2'] and got[0][1] == ['l0', 'l1', 'l2']


def _stub(tf, st):
    return torch.tensor([st.B, st.max_h, st.max_w, tf.size[0], tf.size[1]])


def test_prefetcher_order_and_shapes_with_stubbed_device_half():
    from tpgsr_b200.dataset import BatchPrefetcher
    hr, lr = _crops(10, 1), _crops(10, 2)
    samples = [(hr[i], lr[i], f'w{i}') for i in range(10)]
    pf = BatchPrefetcher(samples, 4, size_hr=(128, 32), down_sample_scale=2, mask=True, depth=2, drop_last=False, _launch=_stub)
    out = list(pf)
    assert [o[2] for o in out] == [['w0', 'w1', 'w2', 'w3'], ['w4', 'w5', 'w6', 'w7'], ['w8', 'w9']]
    for k, (a, b, lab) in enumerate(out):
        sl = slice(4 * k, 4 * k + len(lab))
        assert a.tolist() == [len(lab), max(c.shape[0] for c in hr[sl]), max(c.shape[1] for c in hr[sl]), 128, 32]
        assert b.tolist() == [len(lab), max(c.shape[0] for c in lr[sl]), max(c.shape[1] for c in lr[sl]), 64, 16]
    assert len(list(pf)) == 3                                     # re-iterable


def test_prefetcher_early_exit_and_worker_errors():
    from tpgsr_b200.dataset import BatchPrefetcher
    n0 = threading.active_count()

    def endless():
        i = 0
        while True:
            c = _crops(1, i)[0]
            yield c, c, str(i)
            i += 1

    it = iter(BatchPrefetcher(endless(), 2, _launch=_stub))
    assert next(it)[2] == ['0', '1'] and next(it)[2] == ['2', '3']
    it.close()                                                    # consumer goes away: the worker must stop instead of blocking on a full queue
    for _ in range(50):
        if threading.active_count() <= n0:
            break
        threading.Event().wait(0.1)
    assert threading.active_count() <= n0

    def broken():
        yield _crops(1)[0], _crops(1)[0], 'ok'
        yield np.zeros((3, 3), np.uint8), _crops(1)[0], 'bad image'

    with pytest.raises(ValueError):
        list(BatchPrefetcher(broken(), 1, _launch=_stub))


def test_prefetcher_refuses_cpu_without_stub():
    from tpgsr_b200.dataset import BatchPrefetcher
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    with pytest.raises(RuntimeError):
        next(iter(BatchPrefetcher([], 1)))


def test_lmdb_shaped_source_feeds_batches():
    """TextZoom key layout (dataset/dataset.py:104-149) through an in-memory environment: counts, labels filtered by str_filt, images decoded
    from encoded bytes, an undecodable sample replaced by the next one, and the sample iterable grouped by ``batches``."""
    import io
    import numpy as np
    from PIL import Image
    from tpgsr_b200.dataset.loader import batches
    from tpgsr_b200.dataset.lmdb_source import DictEnv, lmdbDataset_real

    def png(h, w, seed):
        buf = io.BytesIO()
        Image.fromarray(np.random.RandomState(seed).randint(0, 255, (h, w, 3)).astype(np.uint8)).save(buf, format='PNG')
        return buf.getvalue()

    data = {b'num-samples': b'5'}
    words = ['Hello!', 'wor-ld', 'A1', 'xyz?', 'Q']
    for i, wd in enumerate(words, 1):
        data[b'label-%09d' % i] = wd.encode()
        data[b'image_hr-%09d' % i] = png(30 + i, 120 + i, i)
        data[b'image_lr-%09d' % i] = png(15 + i, 60 + i, 100 + i)
    data[b'image_lr-%09d' % 3] = b'not an image'
    ds = lmdbDataset_real(DictEnv(data), voc_type='lower')
    assert len(ds) == 5
    hr, lr, lab = ds[0]
    assert lab == 'hello' and hr.size == (121, 31) and lr.size == (61, 16) and hr.mode == 'RGB'
    assert ds[1][2] == 'world' and ds[2][2] == 'xyz'            # sample 3 is undecodable -> the next one
    got = list(batches(ds, 2, drop_last=False))
    assert [len(b[0]) for b in got] == [2, 2, 1] and got[0][2] == ['hello', 'world']
    import pytest
    with pytest.raises(ImportError):
        lmdbDataset_real('/nonexistent/path')                    # a PATH needs the lmdb package, which this image does not ship
