"""Generate tests/golden/*.npz by running the REFERENCE's own Python modules (TEST INFRASTRUCTURE).

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python -m oracle.gen_golden            # rewrites tests/golden/

The reference modules are imported from ``/root/reference`` (read-only) with a stub ``IPython``
(``model/tsrn.py:9``, ``loss/image_loss.py:6`` import it) and ``matplotlib`` (``model/srcnn.py:10``).
Nothing is copied out of the reference except numerical inputs/outputs.

Weights: every reference module is constructed with its own constructor, then its float
state is overwritten by ``oracle.weights.synth_state_dict`` (deterministic by key name), so the
tests can rebuild the identical weights without the reference.  ``tps.*`` buffers are left as
the reference computed them and are stored in the fixture (the CUDA modules must reproduce them).

Each fixture stores inputs, full outputs, and for gradients a fingerprint per parameter
(L2 norm, sum, first 8 values) — the full tensors are compared on the GPU box against the oracle,
which this file's fixtures pin.
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get('TPGSR_REFERENCE', '/root/reference')
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def _import_reference():
    ip = types.ModuleType('IPython')
    ip.embed = lambda *a, **k: None
    sys.modules.setdefault('IPython', ip)
    mpl = types.ModuleType('matplotlib')
    plt = types.ModuleType('matplotlib.pyplot')
    mpl.pyplot = plt
    sys.modules.setdefault('matplotlib', mpl)
    sys.modules.setdefault('matplotlib.pyplot', plt)
    sys.path.insert(0, REF)
    import importlib
    mods = {}
    for name in ('model.tsrn', 'model.stn_head', 'model.tps_spatial_transformer', 'model.crnn.crnn',
                 'model.srcnn', 'model.srresnet', 'loss.image_loss', 'loss.semantic_loss', 'utils.ssim_psnr'):
        mods[name] = importlib.import_module(name)
    return mods


def fingerprint(t: torch.Tensor) -> np.ndarray:
    f = t.detach().double().reshape(-1)
    head = torch.zeros(8, dtype=torch.float64)
    head[: min(8, f.numel())] = f[:8]
    return torch.cat([torch.stack([f.norm(), f.sum()]), head]).numpy()


def load_synth(module: torch.nn.Module, seed: int):
    from oracle.weights import synth_state_dict
    sd = module.state_dict()
    shapes = {k: tuple(v.shape) for k, v in sd.items()}
    new = synth_state_dict(shapes, seed)
    for k, v in new.items():
        assert sd[k].shape == v.shape, (k, sd[k].shape, v.shape)
        sd[k].copy_(v)
    return shapes


def save(name: str, **arrs):
    os.makedirs(OUT, exist_ok=True)
    conv = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        conv[k] = v
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **conv)
    print('wrote', name, {k: getattr(v, 'shape', None) for k, v in conv.items() if not k.startswith('g/')} if len(conv) < 30 else len(conv))


def grads_fp(module, prefix='g/'):
    out = {}
    for k, p in module.named_parameters():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        out[prefix + k] = fingerprint(g)
    return out


def buffers_fp(module, prefix='buf/'):
    return {prefix + k: fingerprint(b.float()) for k, b in module.named_buffers() if not k.startswith('tps.')}


def main():
    torch.set_num_threads(8)
    from oracle.weights import synth_images
    m = _import_reference()
    tsrn_m, crnn_m = m['model.tsrn'], m['model.crnn.crnn']
    ImageLoss = m['loss.image_loss'].ImageLoss
    SemanticLoss = m['loss.semantic_loss'].SemanticLoss
    shapes_all = {}

    # ------------------------------------------------------------------ op-level
    lr, hr = synth_images(3, seed=11)
    # image loss fwd + grad (loss/image_loss.py)
    sr = (hr * 0.7 + 0.1 * torch.from_numpy(np.random.RandomState(5).rand(*hr.shape).astype(np.float32))).requires_grad_(True)
    crit = ImageLoss(gradient=True, loss_weight=[1, 1e-4])
    loss = crit(sr, hr)
    loss.backward()
    crit_ng = ImageLoss(gradient=False, loss_weight=[1, 1e-4])
    gm = m['loss.image_loss'].GradientPriorLoss.gradient_map(torch.ones(1, 3, 4, 5))
    save('image_loss', sr=sr, hr=hr, loss=loss, grad=sr.grad, loss_nograd=crit_ng(sr.detach(), hr), gradmap_ones=gm)

    # semantic loss (loss/semantic_loss.py)
    r = np.random.RandomState(7)
    pred = torch.softmax(torch.from_numpy(r.randn(26, 5, 37).astype(np.float32)) * 2, -1).requires_grad_(True)
    gt = torch.softmax(torch.from_numpy(r.randn(26, 5, 37).astype(np.float32)) * 2, -1)
    sl = SemanticLoss()(pred, gt)
    sl.backward()
    save('semantic_loss', pred=pred, gt=gt, loss=sl, grad=pred.grad)

    # TPS + grid_sample (model/tps_spatial_transformer.py), perturbed control points incl. out-of-range → clamp path
    tps = m['model.tps_spatial_transformer'].TPSSpatialTransformer(output_image_size=(16, 64), num_control_points=20, margins=(0.05, 0.05))
    from oracle.weights import stn_fc2_bias
    base = stn_fc2_bias(20).reshape(1, 20, 2)
    ctrl = (base + 0.08 * torch.from_numpy(r.randn(3, 20, 2).astype(np.float32))).requires_grad_(True)
    out, src = tps(lr, ctrl)
    w = torch.from_numpy(r.randn(*out.shape).astype(np.float32))
    (out * w).sum().backward()
    ident, _ = tps(lr, tps.target_control_points.unsqueeze(0).expand(3, 20, 2))
    save('tps', img=lr, ctrl=ctrl, out=out, src=src, w=w, grad_ctrl=ctrl.grad, out_identity=ident,
         **{'tps.' + k: v for k, v in tps.state_dict().items()})

    # parse_crnn_data (interfaces/base.py:806-829) from LR and from HR, with grad
    def parse_crnn_data(x):
        y = torch.nn.functional.interpolate(x[:, :3, ...], (32, 100), mode='bicubic')
        return 0.299 * y[:, 0:1] + 0.587 * y[:, 1:2] + 0.114 * y[:, 2:3]
    xl = lr.clone().requires_grad_(True)
    xh = hr.clone().requires_grad_(True)
    gl, gh = parse_crnn_data(xl), parse_crnn_data(xh)
    wl = torch.from_numpy(r.randn(*gl.shape).astype(np.float32))
    ((gl + gh) * wl).sum().backward()
    save('parse_crnn_data', lr=lr, hr=hr, gray_lr=gl, gray_hr=gh, w=wl, grad_lr=xl.grad, grad_hr=xh.grad)

    # PSNR / SSIM (utils/ssim_psnr.py)
    sp = m['utils.ssim_psnr']
    a = hr.clone()
    b = (hr * 0.9 + 0.05).clamp(0, 1)
    save('psnr_ssim', a=a, b=b, psnr=sp.calculate_psnr(a, b), ssim=sp.SSIM()(a, b))

    # GruBlock and one SRB (model/tsrn.py:373-426,491-508)
    torch.manual_seed(0)
    blk = tsrn_m.RecurrentResidualBlock(64).train()
    shapes_all['srb'] = load_synth(blk, 3)
    x = torch.from_numpy(r.randn(2, 64, 16, 64).astype(np.float32) * 0.5).requires_grad_(True)
    y = blk(x)
    wy = torch.from_numpy(r.randn(*y.shape).astype(np.float32))
    (y * wy).sum().backward()
    save('srb', x=x, y=y, w=wy, grad_x=x.grad, **grads_fp(blk), **buffers_fp(blk))
    blk_tl = tsrn_m.RecurrentResidualBlockTL(64, 32).train()
    shapes_all['srb_tl'] = load_synth(blk_tl, 4)
    x2 = x.detach().clone().requires_grad_(True)
    te = torch.from_numpy(r.rand(2, 32, 16, 64).astype(np.float32)).requires_grad_(True)
    y2 = blk_tl(x2, te)
    (y2 * wy).sum().backward()
    save('srb_tl', x=x2, text=te, y=y2, w=wy, grad_x=x2.grad, grad_text=te.grad, **grads_fp(blk_tl))

    # STN head (model/stn_head.py)
    head = m['model.stn_head'].STNHead(in_planes=4, num_ctrlpoints=20, activation='none').train()
    feat0, ctrl0 = head(lr)      # reference init → constant control points KAT
    shapes_all['stn_head'] = load_synth(head, 5)
    xin = lr.clone()
    feat, c = head(xin)
    wc = torch.from_numpy(r.randn(*c.shape).astype(np.float32))
    (c * wc).sum().backward()
    save('stn_head', x=lr, ctrl_at_init=ctrl0, feat=feat, ctrl=c, w=wc, **grads_fp(head), **buffers_fp(head))

    # InfoGen + bilinear(align_corners=True) (model/tsrn.py:81-108,196)
    ig = tsrn_m.InfoGen(37, 32).train()
    shapes_all['infogen'] = load_synth(ig, 6)
    prior = torch.softmax(torch.from_numpy(r.randn(3, 37, 1, 26).astype(np.float32)) * 2, 1).requires_grad_(True)
    t = ig(prior)
    ti = torch.nn.functional.interpolate(t, (16, 64), mode='bilinear', align_corners=True)
    wt = torch.from_numpy(r.randn(*ti.shape).astype(np.float32))
    (ti * wt).sum().backward()
    save('infogen', prior=prior, t=t, t_interp=ti, w=wt, grad_prior=prior.grad, **grads_fp(ig))

    # ------------------------------------------------------------------ model-level
    def model_case(name, ctor, kwargs, batch, seed, training, tl=False, data_seed=21):
        torch.manual_seed(0)
        net = ctor(**kwargs)
        shapes_all[name] = load_synth(net, seed)
        net.train(training)
        lr_i, hr_i = synth_images(batch, seed=data_seed, smooth=(data_seed % 2 == 0))
        args = [lr_i]
        extra = {}
        if tl:
            rr = np.random.RandomState(seed + 100)
            pr = torch.softmax(torch.from_numpy(rr.randn(batch, 37, 1, 26).astype(np.float32)) * 3, 1)
            pr[: batch // 4] = 0
            pr.requires_grad_(True)
            args.append(pr)
            extra['prior'] = pr
        with torch.set_grad_enabled(training):
            sr_o = net(*args)
        out = {'lr': lr_i, 'hr': hr_i, 'sr': sr_o, **extra}
        if training:
            lossv = ImageLoss(gradient=True, loss_weight=[1, 1e-4])(sr_o, hr_i) * 100
            lossv.backward()
            out['loss'] = lossv
            out.update(grads_fp(net))
            out.update(buffers_fp(net))
            if tl:
                out['grad_prior'] = pr.grad
        if hasattr(net, 'tps'):
            out.update({'tps.' + k: v for k, v in net.tps.state_dict().items()})
        save(name, **out)
        return net

    kw = dict(scale_factor=2, width=128, height=32, STN=True, mask=True, srb_nums=5, hidden_units=32)
    model_case('tsrn_train', tsrn_m.TSRN, kw, 3, 1, True)
    model_case('tsrn_eval', tsrn_m.TSRN, kw, 2, 1, False, data_seed=22)
    model_case('tsrn_tl_train', tsrn_m.TSRN_TL, kw, 4, 2, True, tl=True)
    model_case('tsrn_tl_eval', tsrn_m.TSRN_TL, kw, 2, 2, False, tl=True, data_seed=22)

    # CRNN (model/crnn/crnn.py) train + eval
    for training in (True, False):
        torch.manual_seed(0)
        net = crnn_m.CRNN(32, 1, 37, 256)
        shapes_all['crnn'] = load_synth(net, 7)
        net.train(training)
        g = torch.from_numpy(np.random.RandomState(31).rand(3, 1, 32, 100).astype(np.float32)).requires_grad_(training)
        with torch.set_grad_enabled(training):
            lg = net(g)
        out = {'gray': g, 'logits': lg}
        if training:
            wl2 = torch.from_numpy(np.random.RandomState(32).randn(*lg.shape).astype(np.float32))
            (lg * wl2).sum().backward()
            out.update(w=wl2, grad_gray=g.grad, **grads_fp(net), **buffers_fp(net))
        save('crnn_train' if training else 'crnn_eval', **out)

    # SRCNN config 1 (model/srcnn.py:109-145), B=4, MSE*100
    torch.manual_seed(0)
    net = m['model.srcnn'].SRCNN(scale_factor=2, width=128, height=32, STN=False)
    shapes_all['srcnn'] = load_synth(net, 8)
    lr_i, hr_i = synth_images(4, seed=41)
    o = net(lr_i[:, :3])
    lossv = torch.nn.MSELoss()(o, hr_i[:, :3]) * 100
    lossv.backward()
    save('srcnn_train', lr=lr_i, hr=hr_i, sr=o, loss=lossv, **grads_fp(net))

    # ------------------------------------------------------------------ step-level: 3 Adam steps, config-2 style, B=4
    torch.manual_seed(0)
    net = tsrn_m.TSRN(**kw).train()
    load_synth(net, 1)
    crit = ImageLoss(gradient=True, loss_weight=[1, 1e-4])
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, betas=(0.5, 0.999))
    losses, norms = [], []
    for it in range(3):
        lr_i, hr_i = synth_images(4, seed=50 + it)
        lossv = crit(net(lr_i), hr_i).mean() * 100
        opt.zero_grad()
        lossv.backward()
        norms.append(float(torch.nn.utils.clip_grad_norm_(net.parameters(), 0.25)))
        opt.step()
        losses.append(float(lossv))
    fin = {'p/' + k: fingerprint(v.float()) for k, v in net.state_dict().items() if not k.startswith('tps.')}
    save('tsrn_steps', losses=np.array(losses), grad_norms=np.array(norms), **fin)

    # ------------------------------------------------------------------ step-level: TPGSR cascade (config 3/4), B=4, stu_iter=2
    import copy
    torch.manual_seed(0)
    sr_net = tsrn_m.TSRN_TL(**kw).train()
    load_synth(sr_net, 2)
    students = []
    for j in range(2):
        s = crnn_m.CRNN(32, 1, 37, 256)
        load_synth(s, 7 + 10 * j)
        students.append(s.train())
    teacher = crnn_m.CRNN(32, 1, 37, 256)
    load_synth(teacher, 7)
    teacher.eval()
    sem = SemanticLoss()
    params = list(sr_net.parameters()) + [q for s in students for q in s.parameters()]
    opt = torch.optim.Adam(params, lr=1e-3, betas=(0.5, 0.999))
    rec = {}
    for it in range(2):
        lr_i, hr_i = synth_images(4, seed=60 + it)
        with torch.no_grad():
            lab_hr = torch.softmax(teacher(parse_crnn_data(hr_i)), -1)
        cascade = lr_i
        l_img = 0.
        l_dis = 0.
        for i in range(2):
            logits = students[i](parse_crnn_data(cascade[:, :3]))
            lv = torch.softmax(logits, -1)
            lvf = lv.permute(1, 0, 2).unsqueeze(1).permute(0, 3, 1, 2)
            l_dis = l_dis + sem(lv, lab_hr) * 100
            drop = torch.ones(4)
            drop[:1] = 0.
            lvf = lvf * drop.view(-1, 1, 1, 1)
            cascade = sr_net(lr_i, lvf)
            l_img = l_img + crit(cascade, hr_i).mean() * 100
        lossv = l_img + l_dis
        opt.zero_grad()
        lossv.backward()
        if it == 0:
            rec.update({'g0/sr.' + k: fingerprint(q.grad) for k, q in sr_net.named_parameters()})
            for j, s in enumerate(students):
                rec.update({f'g0/stu{j}.' + k: fingerprint(q.grad) for k, q in s.named_parameters()})
            rec['sr0'] = cascade.detach().clone()
        gn = float(torch.nn.utils.clip_grad_norm_(sr_net.parameters(), 0.25))
        opt.step()
        rec[f'loss{it}'] = np.array([float(lossv), float(l_img), float(l_dis), gn])
    rec.update({'p/sr.' + k: fingerprint(v.float()) for k, v in sr_net.state_dict().items() if not k.startswith('tps.')})
    for j, s in enumerate(students):
        rec.update({f'p/stu{j}.' + k: fingerprint(v.float()) for k, v in s.state_dict().items()})
    save('tpgsr_steps', **rec)

    with open(os.path.join(OUT, 'shapes.json'), 'w') as f:
        json.dump({k: {n: list(s) for n, s in v.items()} for k, v in shapes_all.items()}, f, indent=0)
    print('done')


if __name__ == '__main__':
    main()
