"""CPU-side checks: the C-ABI library loads and exports every symbol include/tpgsr_b200.h declares, the drop-in modules
expose the reference's state_dict surface, the flat parameter store aliases correctly, and the product never falls
back to CPU or to the oracle."""
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KW = dict(scale_factor=2, width=128, height=32, STN=True, mask=True, srb_nums=5, hidden_units=32)


def test_library_exports_every_declared_symbol():
    from tpgsr_b200 import _lib
    decls = _lib.parse_header()
    assert len(decls) >= 35
    out = subprocess.run(['nm', '-D', '--defined-only', _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r'\b(tpgsr_\w+)\b', out))
    missing = sorted(set(decls) - exported)
    assert not missing, missing
    for name in decls:                       # ctypes signatures were attached at import
        assert getattr(_lib.LIB, name).argtypes is not None


def test_workspace_queries_run_without_gpu():
    from tpgsr_b200 import _lib
    assert _lib.query('tpgsr_conv2d_wgrad_workspace_bytes', 48, 16, 64, 64, 16, 64, 64, 3, 3) > 0
    assert _lib.query('tpgsr_colreduce_workspace_bytes', 49152, 64) > 0
    assert _lib.query('tpgsr_gru_bwd_workspace_bytes', 768) > 0


@pytest.mark.parametrize('cls_name,key', [('TSRN', 'tsrn_train'), ('TSRN_TL', 'tsrn_tl_train')])
def test_state_dict_surface_matches_reference(shapes, cls_name, key):
    from tpgsr_b200.model import tsrn
    net = getattr(tsrn, cls_name)(**KW)
    sd = net.state_dict()
    assert list(sd.keys()) == list(shapes[key].keys())
    for k, shp in shapes[key].items():
        assert list(sd[k].shape) == shp, k
    # reference KAT: STN head initialised to the constant control points (model/stn_head.py:73-90)
    from oracle.weights import stn_fc2_bias
    assert torch.allclose(net.stn_head.stn_fc2.bias, stn_fc2_bias(20)) and float(net.stn_head.stn_fc2.weight.abs().sum()) == 0
    # TPS buffers reproduce the fixture recorded from the reference
    from conftest import load_golden
    g = load_golden('tps')
    for name in ('inverse_kernel', 'target_coordinate_repr', 'target_control_points', 'padding_matrix'):
        assert torch.allclose(getattr(net.tps, name), torch.from_numpy(g['tps.' + name]), rtol=1e-4, atol=2e-4), name
    assert torch.equal(net.tps.reprT, net.tps.target_coordinate_repr.t())


def test_flat_state_aliases_parameters():
    from tpgsr_b200.model.tsrn import TSRN
    net = TSRN(**KW)
    before = {k: v.clone() for k, v in net.state_dict().items()}
    st = net._flat
    st.ensure()
    assert st.is_current()
    for k, v in net.state_dict().items():
        assert torch.equal(v, before[k]), k
    total = sum(p.numel() for p in net.parameters())
    assert st.n == total
    # every parameter is a view of the flat buffer; fused RNN views cover forward+reverse
    for name, p in net.named_parameters():
        off, n = st.offsets[name]
        assert p.data_ptr() == st.flat.data_ptr() + 4 * off and n == p.numel()
    wih = st.views['block2.gru1.gru.weight_ih']
    assert wih.shape == (192, 64)
    assert torch.equal(wih[:96], net.block2.gru1.gru.weight_ih_l0) and torch.equal(wih[96:], net.block2.gru1.gru.weight_ih_l0_reverse)
    whh = st.views['block6.gru2.gru.weight_hh']
    assert whh.shape == (2, 96, 32) and torch.equal(whh[1], net.block6.gru2.gru.weight_hh_l0_reverse)
    # writes through the flat buffer are visible through the module
    st.flat.zero_()
    assert float(net.block1[0].weight.abs().sum()) == 0
    # load_state_dict keeps the aliasing
    net.load_state_dict(before)
    assert st.is_current() and torch.equal(net.block1[0].weight, before['block1.0.weight'])


def test_no_cpu_fallback():
    from tpgsr_b200.loss.image_loss import ImageLoss
    from tpgsr_b200.model.tsrn import TSRN
    net = TSRN(**KW)
    with pytest.raises(RuntimeError, match='CUDA'):
        net(torch.rand(1, 4, 16, 64))
    with pytest.raises(RuntimeError, match='CUDA'):
        ImageLoss()(torch.rand(1, 4, 32, 128), torch.rand(1, 4, 32, 128))


def test_layout_rules_product_never_imports_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'tpgsr_b200')):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                if re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
    assert 'TEST INFRASTRUCTURE' in open(os.path.join(ROOT, 'oracle', '__init__.py')).read()


def test_srcnn_and_srresnet_state_dict_surfaces(shapes):
    """configs 1 and 5: constructor kwargs and state_dict keys / shapes of the drop-in modules equal the reference's."""
    from tpgsr_b200.model.srcnn import SRCNN
    from tpgsr_b200.model.srresnet import SRResNet_TL
    for net, key in ((SRCNN(scale_factor=2, width=128, height=32, STN=False), 'srcnn'),
                     (SRResNet_TL(scale_factor=2, width=128, height=32, STN=False, mask=True), 'srresnet_tl')):
        sd = net.state_dict()
        assert set(sd) == set(shapes[key]), set(sd) ^ set(shapes[key])
        assert all(list(sd[k].shape) == shapes[key][k] for k in sd)
        with pytest.raises(RuntimeError):
            net(torch.zeros(1, net.in_planes, 16, 64))          # CPU tensors: no fallback
    with pytest.raises(NotImplementedError):
        SRResNet_TL(STN=True)


@pytest.mark.parametrize('nprod', [1, 2, 3])
def test_flat_conv_plans_fit_the_sm(nprod):
    """Host-side tiling plans of the padded-flat tcgen05 kernels (no GPU needed): shared memory <= 227 KB, TMEM <= 512 columns, every
    output row covered, for every geometry the TSRN / CRNN / SRResNet paths send through them."""
    import ctypes
    from tpgsr_b200._lib import LIB
    geoms = [  # B, H, W, P, Ci, Co, KH, KW
        (48, 16, 64, 1, 64, 64, 3, 3), (2, 16, 64, 1, 64, 64, 3, 3), (48, 16, 64, 0, 64, 64, 1, 1), (48, 16, 64, 0, 64, 192, 1, 1),
        (48, 16, 64, 0, 192, 64, 1, 1), (48, 16, 64, 1, 64, 256, 3, 3), (48, 16, 64, 1, 64, 96, 3, 3), (48, 32, 128, 4, 64, 36, 9, 1),
        (48, 32, 128, 4, 64, 64, 9, 1), (48, 16, 64, 4, 64, 64, 9, 1), (48, 16, 64, 4, 64, 36, 9, 1), (48, 4, 16, 1, 64, 128, 3, 3),
        (48, 16, 50, 1, 64, 128, 3, 3), (1248, 1, 1, 0, 512, 2048, 1, 1), (48, 1, 1, 0, 512, 40, 1, 1), (384, 16, 64, 1, 64, 64, 3, 3),
        (1, 5, 17, 1, 64, 64, 3, 3),
        # CRNN conv2..5 and the deep STN-head convs: Ci = 64 * (2, 4, 8) channel blocks
        (48, 8, 25, 1, 128, 256, 3, 3), (48, 8, 25, 1, 256, 256, 3, 3), (48, 4, 26, 1, 256, 512, 3, 3), (48, 4, 26, 1, 512, 512, 3, 3),
        (48, 4, 26, 1, 512, 256, 3, 3), (48, 2, 8, 1, 128, 256, 3, 3), (48, 1, 4, 1, 256, 256, 3, 3), (2, 4, 26, 1, 512, 512, 3, 3),
    ]
    for (B, H, W, P, Ci, Co, KH, KW) in geoms:
        out = (ctypes.c_int * 13)()
        smem = LIB.tpgsr_conv2d_flat_plan(B, H, W, P, Ci, Co, KH, KW, KH // 2, KW // 2, nprod, out)
        assert 0 < smem <= 227 * 1024, ((B, H, W, P, Ci, Co, KH, KW), smem)
        MT, n_tile, n_tiles, tm_stride, G, n_groups, a_stages, box_rows, n_box, tb, b_stages, grid_x, tile_outer = list(out)
        if tile_outer == 3:      # weights-stationary plan (64 -> 64, 3..12 taps, three products): {NT, 64, 1, N, ...}; weights + 2 accumulators in TMEM
            NT, N = MT, tm_stride
            assert nprod == 3 and (Ci, Co) == (64, 64) and N in (64, 96) and KH * KW * 32 + 2 * N <= 512 and box_rows == N // 2
            assert n_box <= 24 and n_box * box_rows >= NT * N + (KH - 1) * (W + P) + KW - 1      # the halo covers every tap of every tile
            assert grid_x * NT * N >= B * (H + P) * (W + P)                                       # every flat output row belongs to a tile
            continue
        assert 1 <= MT <= 3 and MT * tm_stride <= 512 and n_tile % 16 == 0 and n_tile <= 128 and tm_stride <= 256
        assert n_tile * n_tiles >= Co and box_rows % 8 == 0
        if tile_outer == 2:      # tall plan (KH x 1 kernels): MT tiles of 16 rows x 8 columns, one 4-D box [16 + KH - 1][8 MT] per plane
            assert KW == 1 and G == KH and n_groups == 1 and n_box == 1 and box_rows == (16 + KH - 1) * 8 * MT and 8 * MT <= 256
            assert grid_x == B * ((H + 15) // 16) * ((W + 8 * MT - 1) // (8 * MT))    # every output pixel belongs to a tile
            continue
        assert box_rows <= 256
        assert n_box * box_rows >= 128 * MT + (G - 1) * (W + P) + (KW - 1)          # the staged window covers every tap of every tile
        assert grid_x * MT * 128 >= B * (H + P) * (W + P)                           # every flat output row belongs to a tile
        assert b_stages >= 1 and a_stages >= 1 and (n_groups == 1 or a_stages >= 2)
        assert bool(LIB.tpgsr_conv2d_flat_supported(B, H, W, P, Ci, Co, KH, KW, KH // 2, KW // 2, nprod))
    # not supported: channel counts that are no multiple of 64, padding the planes' zero rows / columns do not cover
    assert not LIB.tpgsr_conv2d_flat_supported(48, 16, 64, 1, 96, 64, 3, 3, 1, 1, nprod)
    assert not LIB.tpgsr_conv2d_flat_supported(48, 16, 64, 0, 64, 64, 3, 3, 1, 1, nprod)
    assert not LIB.tpgsr_conv2d_flat_supported(48, 16, 64, 2, 64, 64, 3, 3, 3, 1, nprod)
    # 'valid' and 'full' convolutions (CRNN conv6 2x2 and its data gradient) are
    assert LIB.tpgsr_conv2d_flat_supported(48, 2, 27, 0, 512, 512, 2, 2, 0, 0, nprod)
    assert LIB.tpgsr_conv2d_flat_supported(48, 1, 26, 1, 512, 512, 2, 2, 1, 1, nprod)
    assert LIB.tpgsr_conv2d_flat_supported(48, 16, 64, 1, 64, 64, 3, 3, 0, 0, nprod)
    # the bench shape of the dominant conv: one wave of 139 CTAs x 4 tiles of 96 pixels under the weights-stationary plan
    out = (ctypes.c_int * 13)()
    LIB.tpgsr_conv2d_flat_plan(48, 16, 64, 1, 64, 64, 3, 3, 1, 1, nprod, out)
    assert (out[12] == 3 and out[0] == 4 and out[3] == 96 and out[11] == 139) if nprod == 3 else out[12] == 0
    for (B, H, W, P, Co, K) in [(48, 16, 64, 1, 64, 3), (48, 16, 64, 0, 192, 1), (48, 16, 64, 1, 256, 3), (48, 32, 128, 4, 36, 9)]:
        out = (ctypes.c_int * 9)()
        cpad = (Co + 63) // 64 * 64
        kw = 1 if K == 9 else K
        stage = LIB.tpgsr_conv2d_wgrad_flat_plan(B, H, W, P, 64, Co, cpad, K, kw, nprod, out)
        n_tile, n_tiles, kw_groups, taps, x_rows, stages, total_kb, kb_per_split, splits = list(out)
        assert stage > 0 and stages >= 2 and stages * stage + 2048 <= 227 * 1024
        assert taps * n_tile <= 512 and n_tile * n_tiles == cpad and splits * kb_per_split >= total_kb


def test_pack_store_job_table_layout():
    """ops.PackStore builds the device table tpgsr_pack_weights_batched reads: record size, field order and grid bound (host logic only)."""
    import struct
    import torch
    from tpgsr_b200 import ops
    from tpgsr_b200._lib import LIB
    assert LIB.tpgsr_pack_job_bytes() == struct.calcsize('5Q10i') == 80
    st = ops.PackStore()
    w0 = torch.zeros(64, 64, 3, 3)
    outs0 = tuple(torch.zeros(64, 576, dtype=torch.bfloat16) for _ in range(3)) + (None,)
    st.add('bf16', 'a', 0, w0, outs0)
    w1 = torch.zeros(4, 64, 9, 9)
    outs1 = (torch.zeros(48, 576, dtype=torch.bfloat16), torch.zeros(48, 576, dtype=torch.bfloat16))
    st.add('bf16', ('a', 'kwstack'), 2, w1, outs1, (4, 1, 0, 2, 48))
    w2 = torch.zeros(32, 3, 3, 3)
    st.add('fp32', 'a', 1, w2, (torch.zeros(27, 32), None))          # same name in the other dict: a separate job
    st._build()
    assert st._n == 3 and st._max == 64 * 64 * 9
    raw = bytes(st._table.numpy().tobytes())
    recs = [struct.unpack_from('5Q10i', raw, 80 * i) for i in range(3)]
    assert recs[0][0] == w0.data_ptr() and recs[0][1:5] == (outs0[0].data_ptr(), outs0[1].data_ptr(), outs0[2].data_ptr(), 0)
    assert recs[0][5:10] == (0, 64, 64, 3, 3)
    assert recs[1][5:15] == (2, 4, 64, 9, 9, 4, 1, 0, 2, 48) and recs[1][3:5] == (0, 0)
    assert recs[2][5:10] == (1, 32, 3, 3, 3) and recs[2][2] == 0
    # re-registering a key replaces its job and invalidates the table
    st.add('fp32', 'a', 1, w2, (torch.zeros(27, 32), torch.zeros(288, 3)))
    assert st._table is None and len(st.jobs) == 3


def test_flat_arena_one_bucket_for_sr_and_students():
    """Host logic of the multi-GPU bucket: SR model + students live in ONE parameter / gradient arena (one all-reduce, one
    clip+Adam launch pair), the clip range is the SR model, state_dict values survive the move, a detached parameter is detected."""
    from tpgsr_b200.model._flat import FlatArena
    from tpgsr_b200.model.crnn import CRNN
    from tpgsr_b200.model.tsrn import TSRN_TL
    torch.manual_seed(3)
    sr = TSRN_TL(scale_factor=2, width=128, height=32, STN=True, mask=True, srb_nums=5, hidden_units=32)
    studs = [CRNN(32, 1, 37, 256) for _ in range(2)]
    before = [{k: v.clone() for k, v in m.state_dict().items()} for m in [sr] + studs]
    arena = FlatArena([sr] + studs)
    n_sr = sum(p.numel() for p in sr.parameters())
    assert arena.n_clipped == n_sr and arena.flat.numel() == sum(n for _, n in arena.ranges) == arena.gflat.numel()
    base = arena.flat.data_ptr()
    for (off, n), m in zip(arena.ranges, [sr] + studs):
        assert m._flat.is_current() and m._flat.arena is arena
        for p in m.parameters():
            assert base + 4 * off <= p.data_ptr() < base + 4 * (off + n)
    for m, sd in zip([sr] + studs, before):
        for k, v in m.state_dict().items():
            assert torch.equal(v, sd[k]), k
    # gradients written through a module's views land in the arena's gradient buffer
    studs[1]._flat.gviews['cnn.conv0.weight'].fill_(2.0)
    off, n = arena.ranges[2]
    assert float(arena.gflat[off:off + n].sum()) == 2.0 * studs[1].cnn.conv0.weight.numel()
    # a parameter moved out of the arena (what RNNBase.flatten_parameters does on .to()) is detected: every parameter is checked
    w = studs[0].rnn[0].rnn.weight_hh_l0
    w.data = w.data.clone()
    assert not studs[0]._flat.is_current()
    with pytest.raises(RuntimeError):
        studs[0]._flat.ensure()
    # the per-step guard looks at two sentinels only; flatten_parameters moves every RNN weight, the first in flat order included
    for q in studs[1].rnn[0].rnn.parameters():
        q.data = q.data.clone()
    with pytest.raises(RuntimeError):
        studs[1]._flat.assert_attached()
