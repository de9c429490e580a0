"""GPU parity of the input transform: resizeNormalize for ragged batches through the C ABI, bit-exact against the fixture
recorded from the reference's own class (Pillow bicubic + ToTensor + mean-threshold mask) and against the oracle (oracle/resize.py).
(File name sorts last on purpose: the newest rows run after the hot-path suites.)"""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device('cuda:0')


@pytest.mark.parametrize('tag,size', [('hr', (128, 32)), ('lr', (64, 16))])
@pytest.mark.parametrize('mask', [False, True])
def test_resize_normalize_matches_reference_fixture_bitwise(tag, size, mask):
    from tpgsr_b200.dataset import resizeNormalize
    g = load_golden('resize_normalize')
    crops = [g[f'x/{i}'] for i in range(int(g['n']))]
    want = g[f'y/{tag}/{int(mask)}'].astype(np.float32) / np.float32(255)
    tf = resizeNormalize(size, mask)
    got = tf.batch(crops).cpu().numpy()
    assert got.shape == want.shape and got.dtype == np.float32
    bad = [i for i in range(len(crops)) if not np.array_equal(got[i], want[i])]
    assert not bad, (bad, [crops[i].shape for i in bad])
    one = tf(crops[5]).cpu().numpy()                       # the reference's per-image call
    assert np.array_equal(one, want[5])


def test_resize_normalize_ragged_batch_vs_oracle_and_properties():
    from oracle import resize
    from tpgsr_b200.dataset import resizeNormalize
    r = np.random.RandomState(9)
    crops = [(r.rand(r.randint(1, 80), r.randint(1, 260), 3) * 255).astype(np.uint8) for _ in range(46)]
    crops.append((r.rand(8, 3000, 3) * 255).astype(np.uint8))          # 23x horizontal reduction: weight tables beyond 48 KB of shared memory
    crops.append((r.rand(32, 128, 3) * 255).astype(np.uint8))          # already at the target size: both passes skipped
    tf = resizeNormalize((128, 32), True)
    got = tf.batch(crops).cpu().numpy()
    for i, c in enumerate(crops):
        assert np.array_equal(got[i], resize.resize_normalize(c, (128, 32), True)), (i, c.shape)
    assert np.array_equal(got[-1, :3], crops[-1].transpose(2, 0, 1).astype(np.float32) / np.float32(255))
    assert set(np.unique(got[:, 3])) <= {0.0, 1.0} and got[:, :3].min() >= 0.0 and got[:, :3].max() <= 1.0
    # batch composition does not matter (tables are sized by the largest crop of the batch, results must not depend on it)
    again = tf.batch(crops[:7]).cpu().numpy()
    assert np.array_equal(again, got[:7])


def test_resize_normalize_host_surface():
    from tpgsr_b200 import ops
    from tpgsr_b200._lib import TpgsrError
    from tpgsr_b200.dataset import resizeNormalize
    tf = resizeNormalize((64, 16), False)
    assert tuple(tf.batch([]).shape) == (0, 3, 16, 64)
    with pytest.raises(ValueError):
        tf.batch([np.zeros((4, 4), dtype=np.uint8)])
    with pytest.raises(ValueError):
        tf.batch([np.zeros((4, 4, 3), dtype=np.float32)])
    with pytest.raises(NotImplementedError):
        resizeNormalize((64, 16), False, interpolation=2)
    # C ABI: a scratch buffer that is too small is refused, not overrun
    src = torch.zeros(64, dtype=torch.uint8, device=_dev())
    meta = torch.zeros(4, dtype=torch.int64, device=_dev())
    out = torch.empty(3 * 16 * 64, device=_dev())
    with pytest.raises(TpgsrError):
        ops._call('tpgsr_resize_normalize', src.data_ptr(), meta.data_ptr(), meta.data_ptr(), meta.data_ptr(), 1, 4, 4, 16, 64, 0, out.data_ptr(),
                  None, 0, ops.stream())


def test_batch_prefetcher_matches_direct_transform_and_feeds_a_train_step():
    """BatchPrefetcher (worker-thread packing, copy-stream H2D + resize, event hand-off) yields exactly what resizeNormalize.batch computes,
    in order, and its tensors can be consumed by kernels on the compute stream right away."""
    from tpgsr_b200.dataset import BatchPrefetcher, resizeNormalize
    r = np.random.RandomState(21)
    hr = [(r.rand(r.randint(8, 70), r.randint(16, 220), 3) * 255).astype(np.uint8) for _ in range(22)]
    lr = [(r.rand(r.randint(4, 40), r.randint(8, 120), 3) * 255).astype(np.uint8) for _ in range(22)]
    samples = [(hr[i], lr[i], f'word{i}') for i in range(22)]
    tf_hr, tf_lr = resizeNormalize((128, 32), True), resizeNormalize((64, 16), True)
    seen = 0
    sums = []
    for k, (bh, bl, labels) in enumerate(BatchPrefetcher(samples, 8, size_hr=(128, 32), down_sample_scale=2, mask=True, drop_last=False)):
        sl = slice(8 * k, 8 * k + len(labels))
        assert labels == [f'word{i}' for i in range(sl.start, sl.start + len(labels))]
        assert tuple(bh.shape) == (len(labels), 4, 32, 128) and tuple(bl.shape) == (len(labels), 4, 16, 64)
        sums.append(bh.sum() + bl.sum())                      # consumed on the compute stream immediately (no explicit sync)
        assert torch.equal(bh, tf_hr.batch(hr[sl])) and torch.equal(bl, tf_lr.batch(lr[sl]))
        seen += len(labels)
    assert seen == 22 and all(torch.isfinite(s) for s in sums)


def test_label_onehot_collate_vs_reference_fixture():
    """tpgsr_label_onehot through tpgsr_b200.dataset.label_onehot_collate: bit-exact against the arrays the reference's collate produced."""
    from conftest import load_golden
    from tpgsr_b200.dataset import label_onehot_collate
    g = load_golden('label_collate')
    for name in ('all', 'short', 'with_empty'):
        strs = [str(s) for s in g[name + '/strs']]
        reb, masks, tics = label_onehot_collate(strs)
        assert reb.is_cuda and reb.dtype == torch.float32 and tuple(reb.shape) == g[name + '/label_rebatches'].shape
        assert np.array_equal(reb.cpu().numpy(), g[name + '/label_rebatches']), name
        assert np.array_equal(masks.numpy(), g[name + '/weighted_masks']) and np.array_equal(tics.numpy(), g[name + '/weighted_tics'])
    big = ['w%d' % i * (1 + i % 5) for i in range(512)]          # a batch large enough for several CTAs
    reb, _, _ = label_onehot_collate(big)
    from oracle.collate import label_collate
    assert np.array_equal(reb.cpu().numpy(), label_collate(big)[0])
