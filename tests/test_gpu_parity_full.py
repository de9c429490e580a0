"""GPU parity at the BENCHED shapes and over trajectories:

* the B=48 TSRN train step bench.py times (config 2) vs the CPU oracle: SR output, loss, every gradient tensor element-wise, grad norm;
* a 3-step trajectory on smooth (well-conditioned) crops with ELEMENT-WISE comparison of every updated weight against the oracle's
  trajectory — the quantity compared is the update p_k - p_0 per element, in units of the learning rate;
* the cascade step with stu_iter=3 (config 4's depth) and B=48 eval-mode PSNR / SSIM to 3 dp.

Tolerances (north_star: 1e-3 relative fp32 on outputs; PSNR/SSIM equal to 3 dp): written next to each assert.  Parameters whose gradient
is mathematically zero (a conv bias feeding a train-mode BatchNorm) receive +-lr noise steps from Adam in the reference as well: they are
checked for boundedness only.  STN-head parameters: d(loss)/d(control points) is ill-conditioned on any input with a binarised mask
channel, so trajectories use a smooth 4th channel, where the head is pinned like every other tensor.
"""
import copy
import json
import os

import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu
KW = dict(scale_factor=2, width=128, height=32, STN=True, mask=True, srb_nums=5, hidden_units=32)
LR_RATE = 1e-3
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')


def _dev():
    return torch.device('cuda:0')


def _note(name, obj):
    """Statistics the tolerances were set from: dropped next to the other GPU-run artefacts (scratch, not asserted on)."""
    try:
        os.makedirs(OUT, exist_ok=True)
        with open(os.path.join(OUT, 'parity_stats.jsonl'), 'a') as f:
            f.write(json.dumps({'test': name, **obj}) + '\n')
    except OSError:
        pass


def _smooth_batch(n, seed):
    from oracle.weights import synth_images
    lr, hr = synth_images(n, seed=seed, smooth=True)
    lr[:, 3] = lr[:, :3].mean(1)
    hr[:, 3] = hr[:, :3].mean(1)
    return lr, hr


def _tsrn_sd(shapes, seed=1):
    from oracle.weights import synth_state_dict
    sd = synth_state_dict(shapes['tsrn_train'], seed)
    for k, v in load_golden('tsrn_train').items():
        if k.startswith('tps.'):
            sd[k] = torch.from_numpy(v)
    return sd


def _is_zero_grad_param(k, g_ref):
    return float(g_ref.norm()) < 2e-3


def _compare_updates(name, p0, ours, ref, grads_ref, steps, frac_tol, elem_tol_lr, skip=()):
    """Element-wise: |(ours - p0) - (ref - p0)| <= elem_tol_lr * lr * steps for all but `frac_tol` of each tensor's elements."""
    stats = []
    for k, r in ref.items():
        if any(s in k for s in skip) or k not in ours:
            continue
        du, dr = (ours[k].cpu() - p0[k]).double(), (r - p0[k]).double()
        bound = 2.05 * LR_RATE * steps                   # Adam's |step| <= lr / (1 - beta1) ~ 2 lr early on
        assert float(du.abs().max()) <= bound, (k, float(du.abs().max()))
        if _is_zero_grad_param(k, grads_ref[k]):
            continue
        bad = float(((du - dr).abs() > elem_tol_lr * LR_RATE * steps).double().mean())
        stats.append((bad, rel_err(du, dr), k))
    stats.sort(reverse=True)
    _note(name, {'worst_bad_fraction': stats[:6], 'worst_rel': sorted(((r, k) for _, r, k in stats), reverse=True)[:6]})
    assert stats[0][0] <= frac_tol, stats[:8]
    return stats


def test_tsrn_b48_train_step_vs_oracle(shapes):
    """The exact shape bench.py --config 2 times: B=48, STN on, one CUDA-graphed step."""
    from oracle import step as ostep
    from tpgsr_b200.model.tsrn import TSRN
    from tpgsr_b200.train import TSRNTrainStep
    sd = _tsrn_sd(shapes)
    lr, hr = _smooth_batch(48, 31)
    p = {k: v.clone() for k, v in sd.items()}
    out = ostep.tsrn_train_step(p, lr, hr, ostep.AdamState())
    net = TSRN(**KW)
    net.load_state_dict(sd, strict=False)
    net = net.to(_dev()).train()
    ts = TSRNTrainStep(net, batch=48, use_graph=True)
    loss = float(ts(lr.to(_dev()), hr.to(_dev())))
    assert abs(loss - float(out['loss'])) < 1e-4 * abs(float(out['loss'])), (loss, float(out['loss']))      # 1e-4 relative
    assert rel_err(ts.sr, out['sr']) < 3e-4, rel_err(ts.sr, out['sr'])                                          # budget 1e-3
    # the largest single difference of the 786 432 output values (range [-1, 1]): 1.0e-3 measured — one sample position of the train-mode STN
    # next to a pixel boundary; it moves between 0.9e-3 and 1.0e-3 with the summation order of the BatchNorm statistics
    assert float((ts.sr.cpu() - out['sr']).abs().max()) < 2e-3
    assert abs(ts.opt.grad_norm - float(out['grad_norm'])) < 5e-3 * float(out['grad_norm']), (ts.opt.grad_norm, float(out['grad_norm']))
    worst = []
    for k, g_ref in out['grads'].items():
        got = net._flat.gviews[k].cpu()
        if _is_zero_grad_param(k, g_ref):
            assert float(got.norm()) < 5e-2, (k, float(got.norm()))
            continue
        worst.append((rel_err(got, g_ref), k))
    worst.sort(reverse=True)
    body = [w for w in worst if 'stn_head.' not in w[1]]
    head = [w for w in worst if 'stn_head.' in w[1]]
    _note('tsrn_b48_step', {'loss': [loss, float(out['loss'])], 'sr_rel': rel_err(ts.sr, out['sr']), 'worst_body': body[:8], 'worst_stn_head': head[:4]})
    assert body[0][0] < 5e-3, body[:8]                     # every gradient tensor outside the STN head, rel. L2
    # the localisation head's gradients all carry the error of ONE upstream vector, d(loss)/d(control points): the bilinear sampler's
    # derivative w.r.t. the sampling position is piecewise constant in the position (pixel differences), so it jumps whenever a sample
    # crosses a pixel boundary; two fp32 CPU implementations of the same math already disagree by 3-4% on it (_fp32_floor below)
    assert head[0][0] < 0.15, head[:4]
    ours = {k: v.detach() for k, v in net.state_dict().items()}
    ref = {k: v for k, v in p.items() if ostep.is_param(k)}
    # element-wise Adam updates of every tensor outside the STN head (whose gradient carries the 5% above: its first Adam step is
    # -lr*sign(g), so ~5% of its small-gradient elements flip; measured 3-5%, bounded here at 15%)
    _compare_updates('tsrn_b48_step_updates', sd, ours, ref, out['grads'], 1, frac_tol=0.02, elem_tol_lr=0.05, skip=('stn_head.',))    # measured: 1 of 96
    _compare_updates('tsrn_b48_step_updates_stn', sd, {k: v for k, v in ours.items() if 'stn_head.' in k},
                     {k: v for k, v in ref.items() if 'stn_head.' in k}, out['grads'], 1, frac_tol=0.15, elem_tol_lr=0.05)


def test_tsrn_three_step_trajectory_elementwise(shapes):
    """3 graphed steps on smooth crops (B=8), every weight compared element-wise with the oracle's trajectory after each step."""
    from oracle import step as ostep
    from tpgsr_b200.model.tsrn import TSRN
    from tpgsr_b200.train import TSRNTrainStep
    sd = _tsrn_sd(shapes)
    p = {k: v.clone() for k, v in sd.items()}
    opt = ostep.AdamState()
    net = TSRN(**KW)
    net.load_state_dict(sd, strict=False)
    net = net.to(_dev()).train()
    ts = TSRNTrainStep(net, batch=8, use_graph=True)
    for it in range(3):
        lr, hr = _smooth_batch(8, 40 + it)
        out = ostep.tsrn_train_step(p, lr, hr, opt)
        loss = float(ts(lr.to(_dev()), hr.to(_dev())))
        assert abs(loss - float(out['loss'])) < (1e-4 if it == 0 else 2e-3) * abs(float(out['loss'])), (it, loss, float(out['loss']))
        assert abs(ts.opt.grad_norm - float(out['grad_norm'])) < (5e-3 if it == 0 else 3e-2) * float(out['grad_norm']), (it, ts.opt.grad_norm)
        ours = {k: v.detach() for k, v in net.state_dict().items()}
        ref = {k: v for k, v in p.items() if ostep.is_param(k)}
        # after step 1 an element's update is -lr*sign(g): equal unless the gradient element is at the noise level; later steps are smooth
        # functions of the gradient history: allow 10% of lr per step on all but a few % of a tensor's elements.  The STN head is compared
        # after the first step only: from the second step on its trajectory is chaotic even between two fp32 CPU implementations of the
        # same math (oracle vs stock torch modules, 1e-6 apart: 35% of its elements differ by > 0.1 lr per step after 3 steps, while
        # every other tensor stays below 1% — measured with oracle/torch_eager.py)
        _compare_updates(f'tsrn_traj_step{it}', sd, ours, ref, out['grads'], it + 1, frac_tol=0.02 if it == 0 else 0.05 * it,
                         elem_tol_lr=0.10, skip=() if it == 0 else ('stn_head.',))
        for k, v in net.state_dict().items():
            if 'running_' in k:
                # measured after the third step: 6e-4 outside the STN head, 2e-3 .. 1.2e-2 inside it (its inputs ride the chaotic head trajectory)
                assert rel_err(v, p[k]) < (2e-4 if it == 0 else (5e-2 if 'stn_head.' in k else 2e-3)), (it, k)


def _cascade_models(shapes, n_students):
    from oracle.weights import synth_state_dict
    from tpgsr_b200.model.crnn import CRNN
    from tpgsr_b200.model.tsrn import TSRN_TL
    sd = synth_state_dict(shapes['tsrn_tl_train'], 2)
    for k, v in load_golden('tsrn_tl_train').items():
        if k.startswith('tps.'):
            sd[k] = torch.from_numpy(v)
    sr = TSRN_TL(**KW)
    sr.load_state_dict(sd, strict=False)
    stu_sd = [synth_state_dict(shapes['crnn'], 7 + 10 * j) for j in range(n_students)]
    students = []
    for s in stu_sd:
        m = CRNN(32, 1, 37, 256)
        m.load_state_dict(s)
        students.append(m.to(_dev()).train())
    tea_sd = synth_state_dict(shapes['crnn'], 7)
    teacher = CRNN(32, 1, 37, 256)
    teacher.load_state_dict(tea_sd)
    return sr.to(_dev()).train(), students, teacher.to(_dev()).eval(), sd, stu_sd, tea_sd


def _fp32_floor(sd, stu_sd, tea_sd, lr, hr, stu_iter, out):
    """How far two independent fp32 CPU implementations of the same cascade step (oracle: written-out math; oracle/torch_eager.py: stock
    torch modules) disagree on every gradient tensor: the conditioning of the reference computation itself.  Max-pool arg-max ties, ReLU
    kinks and the sampler's pixel-boundary crossings make the cascade's gradients jump under 1e-7 perturbations."""
    from oracle import torch_eager as te
    sr = te.EagerTSRN(tl=True)
    sr.load_state_dict(sd, strict=True)
    studs = []
    for s_ in stu_sd:
        m = te.EagerCRNN()
        m.load_state_dict(s_, strict=True)
        studs.append(m.train())
    tea = te.EagerCRNN()
    tea.load_state_dict(tea_sd, strict=True)
    step = te.EagerStep(sr.train(), studs, tea.eval(), stu_iter)
    grads = {}
    orig = torch.nn.utils.clip_grad_norm_

    def capture(params, max_norm):
        for k, q in sr.named_parameters():
            grads['sr.' + k] = q.grad.clone()
        for j, m in enumerate(studs):
            for k, q in m.named_parameters():
                grads[f'stu{j}.' + k] = q.grad.clone()
        return orig(params, max_norm)
    torch.nn.utils.clip_grad_norm_ = capture
    try:
        step(lr, hr)
    finally:
        torch.nn.utils.clip_grad_norm_ = orig
    floor = {}
    for k, g in out['sr_grads'].items():
        floor['sr.' + k] = rel_err(grads['sr.' + k], g)
    for j, gd in enumerate(out['stu_grads']):
        for k, g in gd.items():
            floor[f'stu{j}.' + k] = rel_err(grads[f'stu{j}.' + k], g)
    return floor


def test_cascade_stu_iter3_step_vs_oracle(shapes):
    """Config 4's depth: three stages, three students, shared SR model, one arena / one Adam launch; B=8 smooth crops."""
    from oracle import step as ostep
    from tpgsr_b200.train import TPGSRTrainStep
    sr, students, teacher, sd, stu_sd, tea_sd = _cascade_models(shapes, 3)
    lr, hr = _smooth_batch(8, 61)
    p_sr = {k: v.clone() for k, v in sd.items()}
    p_stu = [{k: v.clone() for k, v in s.items()} for s in stu_sd]
    out = ostep.tpgsr_train_step(p_sr, p_stu, {k: v.clone() for k, v in tea_sd.items()}, lr, hr, ostep.AdamState(), stu_iter=3)
    ts = TPGSRTrainStep(sr, students, teacher, batch=8, stu_iter=3, use_distill=True, use_graph=True)
    loss = float(ts(lr.to(_dev()), hr.to(_dev())))
    assert abs(loss - float(out['loss'])) < 2e-4 * abs(float(out['loss'])), (loss, float(out['loss']))
    assert abs(float(ts.loss_img) - float(out['loss_img'])) < 2e-4 * abs(float(out['loss_img']))
    assert abs(float(ts.loss_distill) - float(out['loss_distill'])) < 1e-3 * abs(float(out['loss_distill']))
    assert rel_err(ts.sr, out['sr']) < 5e-4, rel_err(ts.sr, out['sr'])
    assert abs(ts.opt.grad_norm - float(out['grad_norm'])) < 1e-2 * float(out['grad_norm']), (ts.opt.grad_norm, float(out['grad_norm']))
    worst = []
    for k, g_ref in out['sr_grads'].items():
        if _is_zero_grad_param(k, g_ref):
            continue
        worst.append((rel_err(sr._flat.gviews[k], g_ref), 'sr.' + k))
    for j, gd in enumerate(out['stu_grads']):
        for k, g_ref in gd.items():
            if _is_zero_grad_param(k, g_ref):
                continue
            worst.append((rel_err(students[j]._flat.gviews[k], g_ref), f'stu{j}.' + k))
    worst.sort(reverse=True)
    # every gradient tensor within 5x the disagreement of two fp32 CPU implementations on the SAME tensor + 2e-2: three chained
    # CRNN + SR stages are ill-conditioned in the reference itself (measured here, on this batch, not assumed: the two fp32 CPU
    # implementations disagree by 0.6% on the median tensor and 3-4% on the worst; the CUDA path sits at 1.4-1.6% median and 6.5-9% on the
    # worst tensors — always the first student's recurrent layers — and at the fp32 floor when its contractions run in fp32 on the CUDA
    # cores, TPGSR_CONV_MODE=fp32, profiles/r02/README.md).  The worst tensors move by +-1.5 points with ANY change of fp32 summation order:
    # swapping the LSTM / GRU / conv kernels for their other implementations gives 7.5-9.0% on the same tensors
    # (profiles/r02/parity_probe_stu3.log), so the factor is 5, not the 3 that one sample (6.5%) first suggested.
    floor = _fp32_floor(sd, stu_sd, tea_sd, lr, hr, 3, out)
    ratio = sorted(((e / (5 * floor[k] + 2e-2), e, floor[k], k) for e, k in worst), reverse=True)
    _note('cascade_stu3', {'worst_grads': worst[:10], 'worst_vs_floor': ratio[:10], 'loss': [loss, float(out['loss'])],
                           'median_err': worst[len(worst) // 2], 'median_floor': sorted(floor.values())[len(floor) // 2]})
    assert ratio[0][0] <= 1.0, ratio[:10]
    assert worst[len(worst) // 2][0] < 3e-2, worst[len(worst) // 2]
    ours = {k: v.detach() for k, v in sr.state_dict().items()}
    _compare_updates('cascade_stu3_sr_updates', sd, ours, {k: v for k, v in p_sr.items() if ostep.is_param(k)}, out['sr_grads'], 1, 0.15, 0.05, skip=('stn_head.',))
    for j in range(3):
        ours = {k: v.detach() for k, v in students[j].state_dict().items()}
        _compare_updates(f'cascade_stu3_stu{j}_updates', stu_sd[j], ours, {k: v for k, v in p_stu[j].items() if ostep.is_param(k)},
                         out['stu_grads'][j], 1, 0.15, 0.05)          # gradients 6% off (conditioning, see above): up to 9% of a tensor's first Adam steps flip sign
    assert int(sr.block2.bn1.num_batches_tracked) == 3 and int(students[2].cnn.batchnorm2.num_batches_tracked) == 1


def test_eval_b48_psnr_ssim_3dp(shapes):
    """B=48 eval-mode SR of TSRN and of the cascade (stu_iter=1) vs the oracle: 1e-3 relative, PSNR / SSIM equal to 3 dp."""
    from oracle import models, ops as o, step as ostep
    from oracle.weights import synth_images
    from tpgsr_b200.model.tsrn import TSRN
    from tpgsr_b200.train import tpgsr_infer
    lr, hr = synth_images(48, seed=90)
    sd = _tsrn_sd(shapes)
    net = TSRN(**KW)
    net.load_state_dict(sd, strict=False)
    net = net.to(_dev()).eval()
    with torch.no_grad():
        ours = net(lr.to(_dev())).cpu()
        ref = models.tsrn(lr, {k: v.clone() for k, v in sd.items()}, training=False, stn=True)
    assert rel_err(ours, ref) < 2e-4
    assert abs(float(o.calculate_psnr(ours, hr)) - float(o.calculate_psnr(ref, hr))) < 5e-4
    assert abs(float(o.ssim(ours, hr)) - float(o.ssim(ref, hr))) < 5e-4
    sr, students, _, sd_tl, stu_sd, _ = _cascade_models(shapes, 1)
    sr.eval()
    students[0].eval()
    ours = tpgsr_infer(sr, students, lr.to(_dev()), stu_iter=1).cpu()
    ref = ostep.tpgsr_infer({k: v.clone() for k, v in sd_tl.items()}, [dict(stu_sd[0])], lr, stu_iter=1)
    assert rel_err(ours, ref) < 3e-4
    assert abs(float(o.calculate_psnr(ours, hr)) - float(o.calculate_psnr(ref, hr))) < 5e-4
    assert abs(float(o.ssim(ours, hr)) - float(o.ssim(ref, hr))) < 5e-4


def test_config5_srresnet_cascade_step_vs_oracle(shapes):
    """BASELINE config 5's train step (SRResNet_TL + CRNN student / teacher prior, MSE*100, clip on the SR model, Adam over both) through
    TPGSRTrainStep vs the CPU oracle: losses, SR output, every gradient tensor, the element-wise Adam update."""
    from oracle import step as ostep
    from oracle.weights import synth_state_dict
    from tpgsr_b200.model.crnn import CRNN
    from tpgsr_b200.model.srresnet import SRResNet_TL
    from tpgsr_b200.train import TPGSRTrainStep
    sd = synth_state_dict(shapes['srresnet_tl'], 9)
    stu_sd, tea_sd = synth_state_dict(shapes['crnn'], 7), synth_state_dict(shapes['crnn'], 17)
    sr = SRResNet_TL(scale_factor=2, width=128, height=32, STN=False, mask=True)
    sr.load_state_dict(sd)
    stu, tea = CRNN(32, 1, 37, 256), CRNN(32, 1, 37, 256)
    stu.load_state_dict(stu_sd)
    tea.load_state_dict(tea_sd)
    sr, stu, tea = sr.to(_dev()).train(), stu.to(_dev()).train(), tea.to(_dev()).eval()
    lr, hr = _smooth_batch(8, 83)
    p_sr, p_stu = {k: v.clone() for k, v in sd.items()}, {k: v.clone() for k, v in stu_sd.items()}
    out = ostep.tpgsr_train_step(p_sr, [p_stu], dict(tea_sd), lr, hr, ostep.AdamState(), stu_iter=1, gradient=False, arch='srresnet_tl')
    ts = TPGSRTrainStep(sr, [stu], tea, batch=8, stu_iter=1, use_distill=True, gradient=False, use_graph=True)
    loss = float(ts(lr.to(_dev()), hr.to(_dev())))
    assert abs(loss - float(out['loss'])) < 2e-4 * abs(float(out['loss'])), (loss, float(out['loss']))
    assert abs(float(ts.loss_distill) - float(out['loss_distill'])) < 1e-3 * abs(float(out['loss_distill']))
    assert rel_err(ts.sr, out['sr']) < 5e-4, rel_err(ts.sr, out['sr'])
    assert abs(ts.opt.grad_norm - float(out['grad_norm'])) < 1e-2 * float(out['grad_norm']), (ts.opt.grad_norm, float(out['grad_norm']))
    worst = []
    for k, g_ref in out['sr_grads'].items():
        if not _is_zero_grad_param(k, g_ref):
            worst.append((rel_err(sr._flat.gviews[k], g_ref), 'sr.' + k))
    for k, g_ref in out['stu_grads'][0].items():
        if not _is_zero_grad_param(k, g_ref):
            worst.append((rel_err(stu._flat.gviews[k], g_ref), 'stu.' + k))
    worst.sort(reverse=True)
    _note('config5_step', {'worst_grads': worst[:10], 'median': worst[len(worst) // 2], 'loss': [loss, float(out['loss'])]})
    assert worst[0][0] < 5e-2 and worst[len(worst) // 2][0] < 1e-2, worst[:10]      # no STN here; the CRNN's max-pool / ReLU kinks remain
    _compare_updates('config5_sr_updates', sd, {k: v.detach() for k, v in sr.state_dict().items()},
                     {k: v for k, v in p_sr.items() if ostep.is_param(k)}, out['sr_grads'], 1, 0.05, 0.05)
    _compare_updates('config5_stu_updates', stu_sd, {k: v.detach() for k, v in stu.state_dict().items()},
                     {k: v for k, v in p_stu.items() if ostep.is_param(k)}, out['stu_grads'][0], 1, 0.10, 0.05)      # measured: 5.7% of cnn.conv5.bias
