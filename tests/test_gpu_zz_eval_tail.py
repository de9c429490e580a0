"""GPU parity of the eval tail: CTC greedy decode through the C ABI and through the reference-named host function,
bit-exact against the oracle (oracle/ctc.py) and against the strings the reference's own get_string_crnn produced (tests/golden/ctc_decode.npz).
(File name sorts last on purpose: the newest rows run after the hot-path suites.)"""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device('cuda:0')


@pytest.mark.parametrize('case', ['crnn', 'noise', 'long', 't1', 'edge'])
def test_ctc_decode_matches_reference_strings_and_oracle_ids(case):
    from oracle import ctc
    from tpgsr_b200.utils.metrics import ctc_greedy_ids, get_string_crnn
    g = load_golden('ctc_decode')
    x = g['x/' + case]
    xd = torch.from_numpy(x).to(_dev())
    assert get_string_crnn(xd) == [str(s) for s in g['s/' + case]]
    ids, lens = ctc_greedy_ids(xd)
    want_ids, want_lens = ctc.greedy_ids(x)
    assert np.array_equal(lens.cpu().numpy(), want_lens)
    assert np.array_equal(ids.cpu().numpy(), want_ids)          # including the -1 padding


@pytest.mark.parametrize('T,B,C', [(26, 48, 37), (1, 1, 2), (33, 3, 5), (64, 2, 1), (100, 17, 95), (300, 4, 40)])
def test_ctc_decode_shapes_and_properties(T, B, C):
    """Seeded shapes around the warp-chunk boundaries; plus the size-independent properties of a CTC collapse."""
    from oracle import ctc
    from tpgsr_b200.utils.metrics import ctc_greedy_ids
    r = np.random.RandomState(T * 1000 + B * 10 + C)
    x = r.randn(T, B, C).astype(np.float32)
    x[r.rand(T, B) < 0.4, 0] += 4.0                               # plenty of blanks
    x[1:][r.rand(T - 1, B) < 0.4] = 0                             # ... then runs: copy the previous step (done below)
    for t in range(1, T):
        same = (x[t] == 0).all(axis=1)
        x[t, same] = x[t - 1, same]
    ids, lens = ctc_greedy_ids(torch.from_numpy(x).to(_dev()))
    ids, lens = ids.cpu().numpy(), lens.cpu().numpy()
    want_ids, want_lens = ctc.greedy_ids(x)
    assert np.array_equal(lens, want_lens) and np.array_equal(ids, want_ids)
    for b in range(B):
        seq = ids[b, :lens[b]]
        assert (seq > 0).all() and (seq < max(C, 1)).all()          # blanks never decoded
        assert (ids[b, lens[b]:] == -1).all()
    # idempotence: decoding the one-hot logits of (decoded ids interleaved with blanks) gives the same ids back
    T2 = 2 * max(int(lens.max()), 1)
    y = np.full((T2, B, max(C, 2)), -1.0, dtype=np.float32)
    for b in range(B):
        y[:, b, 0] = 1.0
        for k in range(lens[b]):
            y[2 * k, b, 0] = -1.0
            y[2 * k, b, ids[b, k]] = 1.0
    ids2, lens2 = ctc_greedy_ids(torch.from_numpy(y).to(_dev()))
    assert np.array_equal(lens2.cpu().numpy(), lens)
    ids2 = ids2.cpu().numpy()
    for b in range(B):
        assert np.array_equal(ids2[b, :lens[b]], ids[b, :lens[b]]), b


def test_ctc_decode_host_surface():
    from tpgsr_b200.utils.metrics import get_string_crnn
    from tpgsr_b200.utils.util import str_filt
    x = torch.zeros(5, 2, 37, device=_dev())
    x[:, 1, 11] = 1.0
    assert get_string_crnn(x) == ['', 'a']
    assert get_string_crnn(x.permute(1, 0, 2).contiguous().permute(1, 0, 2)) == ['', 'a']       # non-contiguous view
    assert get_string_crnn(x[:, :0]) == []
    with pytest.raises(RuntimeError):
        get_string_crnn(x.cpu())
    with pytest.raises(ValueError):
        get_string_crnn(x, alphabet='-aab' + 'c' * 33)
    assert str_filt('Hello, World!', 'lower') == 'helloworld'


def test_psnr_ssim_cache_is_not_fooled_by_recycled_addresses():
    """cal_psnr(sr, hr); cal_ssim(sr, hr) share one launch, but the next batch — new tensors that the caching allocator puts on the same
    addresses — must be recomputed."""
    from tpgsr_b200.utils.ssim_psnr import SSIM, calculate_psnr
    g = torch.Generator(device='cpu').manual_seed(1)
    seen = []
    for it in range(3):
        a = torch.rand(4, 4, 32, 128, generator=g).to(_dev())
        b = (a * (0.9 - 0.2 * it) + 0.03).clamp(0, 1)
        p, s = float(calculate_psnr(a, b)), float(SSIM()(a, b))
        mse = float((((a[:, :3] * 255) - (b[:, :3] * 255)) ** 2).mean())
        assert abs(p - 20 * np.log10(255.0 / np.sqrt(mse))) < 1e-2, (it, p)
        seen.append((p, s))
        del a, b
    assert len({round(p, 3) for p, _ in seen}) == 3 and len({round(s, 5) for _, s in seen}) == 3
