"""GPU parity, step level, configs 3/4: the TPGSR cascade (students + teacher CRNN, shared TSRN_TL, distillation) through
(a) the fused TPGSRTrainStep and (b) the reference-style loop over the drop-in nn.Modules, vs the fixture recorded from the
reference's own modules (tests/golden/tpgsr_steps.npz) and the CPU oracle."""
import numpy as np
import pytest
import torch

from conftest import assert_fp_close, fingerprint, load_golden, rel_err

pytestmark = pytest.mark.gpu
KW = dict(scale_factor=2, width=128, height=32, STN=True, mask=True, srb_nums=5, hidden_units=32)
T = torch.from_numpy


def _dev():
    return torch.device('cuda:0')


def _build(shapes):
    from oracle.weights import synth_state_dict
    from tpgsr_b200.model.crnn import CRNN
    from tpgsr_b200.model.tsrn import TSRN_TL
    tps = load_golden('tsrn_tl_train')
    sd = synth_state_dict(shapes['tsrn_tl_train'], 2)
    for k, v in tps.items():
        if k.startswith('tps.'):
            sd[k] = T(v)
    sr = TSRN_TL(**KW)
    sr.load_state_dict(sd, strict=False)
    students = []
    for j in range(2):
        s = CRNN(32, 1, 37, 256)
        s.load_state_dict(synth_state_dict(shapes['crnn'], 7 + 10 * j))
        students.append(s.to(_dev()).train())
    teacher = CRNN(32, 1, 37, 256)
    teacher.load_state_dict(synth_state_dict(shapes['crnn'], 7))
    return sr.to(_dev()).train(), students, teacher.to(_dev()).eval()


@pytest.mark.parametrize('use_graph', [False, True])
def test_fused_cascade_steps_vs_reference_fixture(shapes, use_graph):
    from oracle.weights import synth_images
    from tpgsr_b200.train import TPGSRTrainStep
    g = load_golden('tpgsr_steps')
    sr, students, teacher = _build(shapes)
    ts = TPGSRTrainStep(sr, students, teacher, batch=4, stu_iter=2, use_distill=True, use_graph=use_graph)
    for it in range(2):
        lr, hr = synth_images(4, seed=60 + it)
        loss = float(ts(lr.to(_dev()), hr.to(_dev())))
        want = g[f'loss{it}']
        lt = 1e-3 if it == 0 else 1e-2
        assert abs(loss - want[0]) < lt * abs(want[0]), (it, loss, want)
        assert abs(float(ts.loss_img) - want[1]) < lt * abs(want[1])
        assert abs(float(ts.loss_distill) - want[2]) < 4 * lt * abs(want[2])
        assert abs(ts.opt_sr.grad_norm - want[3]) < (3e-2 if it == 0 else 0.6) * want[3], (ts.opt_sr.grad_norm, want[3])
        if it == 0:
            assert rel_err(ts.sr, g['sr0']) < 1e-3
    for k, v in sr.state_dict().items():
        if k.startswith('tps.') or 'stn_head.' in k or k.endswith('num_batches_tracked'):
            continue
        assert_fp_close(fingerprint(v.float()), g['p/sr.' + k], k, rtol=1e-2, atol=1e-2, noise_floor=0.0, check_head=False)
    assert int(sr.block2.bn1.num_batches_tracked) == 4          # 2 steps x 2 cascade stages through the shared SR model
    assert int(students[0].cnn.batchnorm2.num_batches_tracked) == 2


def test_module_api_cascade_matches_fixture_first_step(shapes):
    """The reference's own loop shape (interfaces/super_resolution.py:297-424) over the drop-in modules + torch glue."""
    from oracle.weights import synth_images
    from tpgsr_b200.loss.image_loss import ImageLoss
    from tpgsr_b200.loss.semantic_loss import SemanticLoss
    g = load_golden('tpgsr_steps')
    sr, students, teacher = _build(shapes)
    crit, sem = ImageLoss(gradient=True, loss_weight=[1, 1e-4]), SemanticLoss()

    def parse_crnn_data(x):
        y = torch.nn.functional.interpolate(x[:, :3, ...], (32, 100), mode='bicubic')
        return 0.299 * y[:, 0:1] + 0.587 * y[:, 1:2] + 0.114 * y[:, 2:3]

    lr, hr = synth_images(4, seed=60)
    lr, hr = lr.to(_dev()), hr.to(_dev())
    with torch.no_grad():
        lab_hr = torch.softmax(teacher(parse_crnn_data(hr)), -1)
    cascade, l_img, l_dis = lr, 0., 0.
    for i in range(2):
        logits = students[i](parse_crnn_data(cascade[:, :3]))
        lv = torch.softmax(logits, -1)
        lvf = lv.permute(1, 0, 2).unsqueeze(1).permute(0, 3, 1, 2)
        l_dis = l_dis + sem(lv, lab_hr) * 100
        drop = torch.ones(4, device=_dev())
        drop[:1] = 0.
        cascade = sr(lr, lvf * drop.view(-1, 1, 1, 1))
        l_img = l_img + crit(cascade, hr).mean() * 100
    loss = l_img + l_dis
    loss.backward()
    want = g['loss0']
    assert abs(float(loss) - want[0]) < 1e-3 * abs(want[0])
    gn = float(torch.nn.utils.clip_grad_norm_(sr.parameters(), 0.25))
    assert abs(gn - want[3]) < 3e-2 * want[3], (gn, want[3])
    for j, s in enumerate(students):
        for k, q in s.named_parameters():
            assert_fp_close(fingerprint(q.grad), g[f'g0/stu{j}.' + k], k, rtol=8e-2, atol=2e-3, check_head=False)


def test_cascade_inference_vs_oracle(shapes):
    from oracle import step
    from oracle.weights import synth_images, synth_state_dict
    from tpgsr_b200.train import tpgsr_infer
    sr, students, _ = _build(shapes)
    sr.eval()
    for s in students:
        s.eval()
    lr, _ = synth_images(3, seed=71)
    out = tpgsr_infer(sr, students, lr.to(_dev()), stu_iter=2)
    tps = load_golden('tsrn_tl_train')
    sr_p = synth_state_dict(shapes['tsrn_tl_train'], 2)
    for k, v in tps.items():
        if k.startswith('tps.'):
            sr_p[k] = T(v)
    stu_p = [synth_state_dict(shapes['crnn'], 7 + 10 * j) for j in range(2)]
    ref = step.tpgsr_infer(sr_p, stu_p, lr, stu_iter=2)
    assert rel_err(out, ref) < 3e-4
    from oracle import ops as o
    _, hr = synth_images(3, seed=71)
    assert abs(float(o.calculate_psnr(out.cpu(), hr)) - float(o.calculate_psnr(ref, hr))) < 5e-4
    assert abs(float(o.ssim(out.cpu(), hr)) - float(o.ssim(ref, hr))) < 5e-4
