"""GPU parity tests, module / step level: the drop-in nn.Modules (tpgsr_b200.model.*) through their public
forward()/backward() against the CPU oracle on identical seeded weights and inputs, and against the golden fixtures
recorded from the reference's own modules."""
import copy

import numpy as np
import pytest
import torch

from conftest import assert_fp_close, fingerprint, load_golden, rel_err

pytestmark = pytest.mark.gpu

KW = dict(scale_factor=2, width=128, height=32, STN=True, mask=True, srb_nums=5, hidden_units=32)


def _dev():
    return torch.device('cuda:0')


def _load(module, sd):
    missing, unexpected = module.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all('reprT' in m or m.startswith('tps.') for m in missing), missing
    return module


def _synth(shapes, key, seed, golden=None):
    from oracle.weights import synth_state_dict
    sd = synth_state_dict(shapes[key], seed)
    if golden is not None:
        for k, v in golden.items():
            if k.startswith('tps.'):
                sd[k] = torch.from_numpy(v)
    return sd


def _grad_report(named_params, oracle_grads, tol, skip=()):
    worst = []
    for k, p in named_params:
        if any(s in k for s in skip):
            continue
        want = oracle_grads[k]
        got = p.grad if p.grad is not None else torch.zeros_like(p)
        if float(want.norm()) < 2e-3:          # mathematically-zero grads (bias before train-mode BN)
            assert float(got.norm()) < 5e-2, (k, float(got.norm()))
            continue
        worst.append((rel_err(got, want), k))
    worst.sort(reverse=True)
    assert worst[0][0] < tol, worst[:8]
    return worst[0]


@pytest.mark.parametrize('smooth', [True, False])
def test_tsrn_train_forward_backward_vs_oracle(shapes, smooth):
    from oracle import models, ops as o, step
    from oracle.weights import synth_images
    from tpgsr_b200.loss.image_loss import ImageLoss
    from tpgsr_b200.model.tsrn import TSRN
    g = load_golden('tsrn_train')
    sd = _synth(shapes, 'tsrn_train', 1, g)
    lr, hr = synth_images(3, seed=20 if smooth else 21, smooth=smooth)
    if smooth:
        # the binarised mask channel is pixel noise, which makes d(loss)/d(sampling grid) chaotic even on blurred RGB;
        # for the tight STN-gradient check use a smooth 4th channel as well
        lr[:, 3] = lr[:, :3].mean(1)
        hr[:, 3] = hr[:, :3].mean(1)
    # oracle
    p = {k: v.clone() for k, v in sd.items()}
    leaves = {k: v.requires_grad_(True) for k, v in p.items() if step.is_param(k)}
    sr_o = models.tsrn(lr, p, True, stn=True)
    loss_o = o.image_loss(sr_o, hr) * 100
    loss_o.backward()
    og = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    # CUDA path through the public nn.Module API
    net = _load(TSRN(**KW), sd).to(_dev()).train()
    crit = ImageLoss(gradient=True, loss_weight=[1, 1e-4])
    sr = net(lr.to(_dev()))
    loss = crit(sr, hr.to(_dev())).mean() * 100
    loss.backward()
    assert sr.shape == (3, 4, 32, 128)
    assert rel_err(sr, sr_o) < (3e-4 if smooth else 1e-3), rel_err(sr, sr_o)
    assert abs(float(loss) - float(loss_o)) < 1e-4 * abs(float(loss_o))
    # d(loss)/d(control points) is ill-conditioned at the STN's operating point: perturbing ctrl by 1e-6 (relative) changes
    # the ORACLE's own autograd gradient by ~4% (scripts/diag_stn.py, /tmp-free restatement in DESIGN.md §parity), and the
    # TPS solve only delivers `src` to ~1e-5.  So the STN-head gradients are pinned piecewise instead: d/d(rectified image)
    # here (block1.* grads), TPS backward in test_gpu_ops.py (9e-5 with the reference's src), the head's own backward in
    # test_stn_head_forward_backward_vs_oracle_and_golden; here they only have to agree to the conditioning level.
    _grad_report(net.named_parameters(), og, 2e-3 if smooth else 3e-2, skip=('stn_head.',))
    stn = [(rel_err(q.grad, og[k]), k) for k, q in net.named_parameters() if 'stn_head.' in k and float(og[k].norm()) > 2e-3]
    assert max(stn)[0] < 0.25, sorted(stn, reverse=True)[:5]
    for k, b in net.named_buffers():
        if 'running_' in k:
            assert rel_err(b, p[k]) < 1e-4, k
        if k.endswith('num_batches_tracked'):
            assert int(b) == 1
    if not smooth:
        # same inputs as the fixture recorded from the reference modules
        assert rel_err(sr, g['sr']) < 5e-4
        assert abs(float(loss) - float(g['loss'])) < 1e-4 * abs(float(g['loss']))
        for k, q in net.named_parameters():
            assert_fp_close(fingerprint(q.grad), g['g/' + k], k, rtol=3e-2, atol=1e-3, check_head=False)


def test_tsrn_eval_vs_golden(shapes):
    from tpgsr_b200.model.tsrn import TSRN
    g = load_golden('tsrn_eval')
    net = _load(TSRN(**KW), _synth(shapes, 'tsrn_train', 1, g)).to(_dev()).eval()
    with torch.no_grad():
        sr = net(torch.from_numpy(g['lr']).to(_dev()))
    assert rel_err(sr, g['sr']) < 5e-5
    # PSNR / SSIM of our output vs the reference output agree to 3 dp with reference-vs-itself (inf / 1.0 excluded):
    from oracle import ops as o
    hr = torch.from_numpy(g['hr'])
    assert abs(float(o.calculate_psnr(sr.cpu(), hr)) - float(o.calculate_psnr(torch.from_numpy(g['sr']), hr))) < 5e-4
    assert abs(float(o.ssim(sr.cpu(), hr)) - float(o.ssim(torch.from_numpy(g['sr']), hr))) < 5e-4


def test_tsrn_state_dict_roundtrip_and_grad_accumulation(shapes):
    from tpgsr_b200.loss.image_loss import ImageLoss
    from tpgsr_b200.model.tsrn import TSRN
    from oracle.weights import synth_images
    sd = _synth(shapes, 'tsrn_train', 1, load_golden('tsrn_train'))
    net = _load(TSRN(**KW), sd).to(_dev()).train()
    out = net.state_dict()
    assert list(out.keys()) == list(shapes['tsrn_train'].keys())
    for k, v in sd.items():
        assert torch.equal(out[k].cpu(), v), k
    lr, hr = synth_images(2, seed=5, smooth=True)
    crit = ImageLoss()
    net.zero_grad()
    (crit(net(lr.to(_dev())), hr.to(_dev())) * 100).backward()
    g1 = {k: p.grad.clone() for k, p in net.named_parameters()}
    # a second backward without zero_grad accumulates (running stats differ slightly, so compare loosely)
    (crit(net(lr.to(_dev())), hr.to(_dev())) * 100).backward()
    for k, p in net.named_parameters():
        if float(g1[k].norm()) > 1e-2:
            assert rel_err(p.grad, 2 * g1[k]) < 1e-3, k


def test_tsrn_three_graphed_steps_vs_golden_and_oracle(shapes):
    """TSRNTrainStep (CUDA-graphed fwd+loss+bwd+clip+Adam) reproduces the reference's 3-step trajectory."""
    from oracle import step
    from oracle.weights import synth_images
    from tpgsr_b200.model.tsrn import TSRN
    from tpgsr_b200.train import TSRNTrainStep
    g = load_golden('tsrn_steps')
    sd = _synth(shapes, 'tsrn_train', 1, load_golden('tsrn_train'))
    for use_graph in (False, True):
        net = _load(TSRN(**KW), copy.deepcopy(sd)).to(_dev()).train()
        ts = TSRNTrainStep(net, batch=4, use_graph=use_graph)
        for it in range(3):
            lr, hr = synth_images(4, seed=50 + it)
            loss = float(ts(lr.to(_dev()), hr.to(_dev())))
            ltol = 3e-4 if it == 0 else 1e-2       # after noisy Adam steps (see tests/test_oracle_golden.py)
            assert abs(loss - g['losses'][it]) < ltol * g['losses'][it], (use_graph, it, loss, g['losses'][it])
            gn = ts.opt.grad_norm
            # STN gradients on white-noise crops are chaotic at the % level; after Adam steps the norm is only a sanity
            # bound (the oracle itself is 10% off the reference at step 2: tests/test_oracle_golden.py)
            tol = 1e-2 if it == 0 else 0.6
            assert abs(gn - g['grad_norms'][it]) < tol * g['grad_norms'][it], (use_graph, it, gn, g['grad_norms'][it])
        for k, v in net.state_dict().items():
            if k.startswith('tps.') or 'stn_head.' in k:
                continue
            if k.endswith('num_batches_tracked'):
                assert int(v) == 3, k
                continue
            assert_fp_close(fingerprint(v.float()), g['p/' + k], k, rtol=5e-3, atol=7e-3, noise_floor=0.0, check_head=False)


def test_stn_head_forward_backward_vs_oracle_and_golden(shapes):
    """The localisation head in isolation (so its backward is pinned without the ill-conditioned grid gradient)."""
    from oracle import models, step
    from tpgsr_b200 import engine
    from tpgsr_b200.model.tsrn import TSRN
    g = load_golden('stn_head')
    from oracle.weights import synth_state_dict
    head_sd = {'stn_head.' + k: v for k, v in synth_state_dict(shapes['stn_head'], 5).items()}
    x = torch.from_numpy(g['x'])
    w = torch.from_numpy(g['w'])
    p = {k: v.clone() for k, v in head_sd.items()}
    leaves = {k: v.requires_grad_(True) for k, v in p.items() if step.is_param(k)}
    feat_o, ctrl_o = models.stn_head(x, p, 'stn_head.', True)
    (ctrl_o * w).sum().backward()
    net = TSRN(**KW)
    net.load_state_dict(head_sd, strict=False)
    net = net.to(_dev()).train()
    st = net._flat
    st.ensure()
    ctx = engine.Ctx(st.views, st.gviews, True)
    xn = x.permute(0, 2, 3, 1).contiguous().to(_dev())
    feat, ctrl, saved = engine.stn_head_fwd(ctx, xn)
    assert rel_err(ctrl, g['ctrl']) < 2e-5 and rel_err(feat, g['feat']) < 2e-4
    engine.stn_head_bwd(ctx, saved, w.to(_dev()))
    worst = []
    for k, v in leaves.items():
        if float(v.grad.norm()) < 2e-3:
            continue
        worst.append((rel_err(st.gviews[k], v.grad), k))
        assert_fp_close(fingerprint(st.gviews[k]), g['g/' + k[len('stn_head.'):]], k, rtol=1e-2, atol=1e-4, check_head=False)
    worst.sort(reverse=True)
    assert worst[0][0] < 5e-3, worst[:5]
    for k, v in p.items():
        if 'running_' in k:
            assert rel_err(st.views[k], v) < 1e-4, k


def test_srcnn_config1_vs_golden_and_oracle(shapes):
    """BASELINE config 1 (SRCNN, B=4, `images[:, :3]`, MSE*100): forward, loss, every parameter gradient vs the fixture recorded
    from the reference module, and the full output vs the oracle."""
    from oracle import models
    from oracle.weights import synth_state_dict
    from tpgsr_b200.model.srcnn import SRCNN
    g = load_golden('srcnn_train')
    sd = synth_state_dict(shapes['srcnn'], 8)
    net = SRCNN(scale_factor=2, width=128, height=32, STN=False)
    net.load_state_dict(sd)
    net = net.to(_dev()).train()
    lr, hr = torch.from_numpy(g['lr']).to(_dev()), torch.from_numpy(g['hr']).to(_dev())
    sr = net(lr[:, :3])
    loss = torch.nn.MSELoss()(sr, hr[:, :3]) * 100
    loss.backward()
    assert rel_err(sr, g['sr']) < 1e-4, rel_err(sr, g['sr'])
    assert rel_err(sr, models.srcnn(torch.from_numpy(g['lr'])[:, :3], sd)) < 1e-4
    assert abs(float(loss) - float(g['loss'])) < 1e-3 * float(g['loss'])
    for k, p in net.named_parameters():
        assert_fp_close(fingerprint(p.grad), g['g/' + k], k, rtol=2e-3)
    with torch.no_grad():
        net.eval()
        sr2 = net(lr[:, :3])
    assert rel_err(sr2, g['sr']) < 1e-4


def test_srresnet_tl_config5_vs_golden_and_oracle(shapes):
    """BASELINE config 5 module (SRResNet_TL + text prior, MSE*100): forward, loss, every parameter gradient and the prior gradient
    vs the fixture recorded from the reference module; eval mode with and without a prior."""
    from oracle.weights import synth_state_dict
    from tpgsr_b200.model.srresnet import SRResNet_TL
    g = load_golden('srresnet_tl')
    sd = synth_state_dict(shapes['srresnet_tl'], 9)
    for k, v in g.items():
        if k.startswith('slope/'):
            sd[k[len('slope/'):]] = torch.from_numpy(v)
    net = SRResNet_TL(scale_factor=2, width=128, height=32, STN=False, mask=True)
    net.load_state_dict(sd)
    net = net.to(_dev()).train()
    lr, hr = torch.from_numpy(g['lr']).to(_dev()), torch.from_numpy(g['hr']).to(_dev())
    prior = torch.from_numpy(g['prior']).to(_dev()).requires_grad_(True)
    sr = net(lr, prior)
    loss = torch.nn.MSELoss()(sr, hr) * 100
    loss.backward()
    assert rel_err(sr, g['sr']) < 2e-4, rel_err(sr, g['sr'])
    assert abs(float(loss) - float(g['loss'])) < 1e-3 * float(g['loss'])
    assert rel_err(prior.grad, g['grad_prior']) < 5e-3, rel_err(prior.grad, g['grad_prior'])
    for k, p in net.named_parameters():
        want = g['g/' + k]
        if p.numel() == 1:       # PReLU slopes: sums of ~10^6 signed terms
            assert abs(float(p.grad) - float(want[1])) < 1e-2 * abs(float(want[1])) + 1e-4, k
        else:
            # norm / sum at 5e-3; the first 8 elements relative to the tensor's norm (train-mode BN on a 4-image batch amplifies fp32
            # summation-order differences: single small elements move by a few 1e-3 of the norm between correct implementations —
            # block4.bn1.bias[2] sits at 3.3e-3 of the norm with the weights-stationary conv plan, whose own error against an fp64 conv
            # is LOWER than the flat plan's: 3.65e-6 vs 4.56e-6, profiles/r02/ws_timeline_v4.log)
            got = fingerprint(p.grad)
            assert_fp_close(got, want, k, rtol=5e-3, check_head=False)
            if abs(want[0]) >= 2e-3:
                assert np.all(np.abs(got[2:] - want[2:]) <= 5e-3 * abs(want[0]) + 2e-5), (k, got[2:], want[2:])
    for k, b in net.named_buffers():
        if 'running_' in k:
            assert_fp_close(fingerprint(b), g['buf/' + k], k, rtol=2e-3)
    with torch.no_grad():
        net.eval()
        assert rel_err(net(lr, prior.detach()), g['sr_eval']) < 2e-4
        assert rel_err(net(lr), g['sr_eval_noprior']) < 2e-4


def test_tsrn_without_mask_channel_vs_oracle():
    """``mask=False`` (3 input planes; ``--mask`` is a CLI flag the reference passes through, interfaces/base.py:296-297): train-mode forward
    + backward and eval forward vs the CPU oracle."""
    from oracle import models, ops as o, step
    from oracle.ops import tps_buffers
    from oracle.weights import synth_images, synth_state_dict
    from tpgsr_b200.loss.image_loss import ImageLoss
    from tpgsr_b200.model.tsrn import TSRN
    kw = dict(KW, mask=False)
    net = TSRN(**kw)
    assert net.in_planes == 3 and net.block1[0].weight.shape[1] == 3 and net.block8[1].weight.shape[0] == 3
    sd = synth_state_dict({k: list(v.shape) for k, v in net.state_dict().items()}, 4)
    sd.update({'tps.' + k: v for k, v in tps_buffers(16, 64).items()})
    net.load_state_dict(sd, strict=False)
    net = net.to(_dev()).train()
    lr, hr = synth_images(3, seed=33, smooth=True)
    lr, hr = lr[:, :3].contiguous(), hr[:, :3].contiguous()
    p = {k: v.clone() for k, v in sd.items()}
    leaves = {k: v.requires_grad_(True) for k, v in p.items() if step.is_param(k)}
    sr_o = models.tsrn(lr, p, True, stn=True)
    loss_o = o.image_loss(sr_o, hr) * 100
    loss_o.backward()
    sr = net(lr.to(_dev()))
    loss = ImageLoss(gradient=True, loss_weight=[1, 1e-4])(sr, hr.to(_dev())).mean() * 100
    loss.backward()
    assert sr.shape == (3, 3, 32, 128) and rel_err(sr, sr_o) < 3e-4, rel_err(sr, sr_o)
    assert abs(float(loss) - float(loss_o)) < 1e-4 * abs(float(loss_o))
    og = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    _grad_report(net.named_parameters(), og, 3e-3, skip=('stn_head.',))
    net.eval()
    with torch.no_grad():
        assert rel_err(net(lr.to(_dev())), models.tsrn(lr, {k: v.detach() for k, v in p.items()}, False, stn=True)) < 1e-4
