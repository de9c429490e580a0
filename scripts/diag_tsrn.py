"""Diagnostic (GPU box): per-parameter gradient error of the CUDA TSRN vs the CPU oracle."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from conftest import load_golden, rel_err
from oracle import models, ops as o, step
from oracle.weights import synth_state_dict, synth_images
from tpgsr_b200.loss.image_loss import ImageLoss
from tpgsr_b200.model.tsrn import TSRN
KW = dict(scale_factor=2, width=128, height=32, STN=True, mask=True, srb_nums=5, hidden_units=32)
shapes = json.load(open(os.path.join(ROOT, 'tests/golden/shapes.json')))
g = load_golden('tsrn_train')
for smooth, seed, B in ((True, 20, 3), (False, 21, 3), (False, 50, 4)):
    sd = synth_state_dict(shapes['tsrn_train'], 1)
    for k, v in g.items():
        if k.startswith('tps.'): sd[k] = torch.from_numpy(v)
    lr, hr = synth_images(B, seed=seed, smooth=smooth)
    p = {k: v.clone() for k, v in sd.items()}
    leaves = {k: v.requires_grad_(True) for k, v in p.items() if step.is_param(k)}
    sr_o, aux = models.tsrn(lr, p, True, stn=True, return_aux=True)
    loss_o = o.image_loss(sr_o, hr) * 100
    loss_o.backward()
    net = TSRN(**KW); net.load_state_dict(sd, strict=False); net = net.cuda().train()
    sr = net(lr.cuda()); loss = ImageLoss(True, [1, 1e-4])(sr, hr.cuda()) * 100; loss.backward()
    print(f'--- smooth={smooth} seed={seed} B={B}: sr rel {rel_err(sr, sr_o):.2e} loss {float(loss):.6f} vs {float(loss_o):.6f}')
    rows = []
    tot_o = tot_g = 0.0
    for k, q in net.named_parameters():
        want = leaves[k].grad if leaves[k].grad is not None else torch.zeros_like(leaves[k])
        rows.append((rel_err(q.grad, want), float(want.norm()), float(q.grad.norm()), k))
        tot_o += float(want.norm()) ** 2; tot_g += float(q.grad.norm()) ** 2
    print('  total grad norm oracle %.4f cuda %.4f' % (tot_o ** 0.5, tot_g ** 0.5))
    for r in sorted(rows, reverse=True)[:12]:
        print('   rel %.2e  |oracle| %.4e |cuda| %.4e  %s' % r)
