"""Diagnostic (GPU box): locate the STN-path gradient discrepancy (dxr -> dctrl -> head grads)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from conftest import load_golden, rel_err
from oracle import models, ops as o, step
from oracle.weights import synth_state_dict, synth_images
from tpgsr_b200 import engine, ops
from tpgsr_b200.model.tsrn import TSRN
KW = dict(scale_factor=2, width=128, height=32, STN=True, mask=True, srb_nums=5, hidden_units=32)
shapes = json.load(open(os.path.join(ROOT, 'tests/golden/shapes.json')))
g = load_golden('tsrn_train')
sd = synth_state_dict(shapes['tsrn_train'], 1)
for k, v in g.items():
    if k.startswith('tps.'): sd[k] = torch.from_numpy(v)
lr, hr = synth_images(3, seed=20, smooth=True)
lr[:, 3] = lr[:, :3].mean(1); hr[:, 3] = hr[:, :3].mean(1)
p = {k: v.clone() for k, v in sd.items()}
leaves = {k: v.requires_grad_(True) for k, v in p.items() if step.is_param(k)}
sr_o, aux = models.tsrn(lr, p, True, stn=True, return_aux=True)
aux['ctrl'].retain_grad(); aux['rectified'].retain_grad()
loss_o = o.image_loss(sr_o, hr) * 100
loss_o.backward()
net = TSRN(**KW); net.load_state_dict(sd, strict=False); net = net.cuda().train()
st = net._flat; st.ensure()
ctx = engine.Ctx(st.views, st.gviews, True)
x = lr.cuda()
sr, saved = engine.tsrn_fwd(ctx, x, True, 5, None)
s_stn, s_b1, b1pre, s_srb, s_c7, s_bn7, s_up, u, s_out, srt, dims = saved
s_head, xn, src = s_stn
nchw = lambda t: t.permute(0, 3, 1, 2).contiguous()
xr = s_b1[0]
print('ctrl fwd rel', rel_err(src, src), 'rectified rel', rel_err(nchw(xr), aux['rectified']))
dsr = ops.image_loss_bwd(sr, hr.cuda(), True, 1.0, 1e-4, 100.0)
# replicate tsrn_bwd up to dxr
dpre = ops.nchw_to_nhwc(dsr, tanh_out=sr)
dv = engine.conv_bwd(ctx, s_out, dpre)
du = ops.pixelshuffle_mish_bwd(dv, u)
ds8 = engine.conv_bwd(ctx, s_up, du)
dc7 = engine.bn_bwd(ctx, s_bn7, ds8)
db = engine.conv_bwd(ctx, s_c7, dc7)
for s in reversed(s_srb):
    db = engine.srb_bwd(ctx, s, db, None)
db1 = ops.add(db, ds8)
gs, acc = ctx.G('block1.1.weight')
db1pre = ops.act_bwd(db1, b1pre, ops.ACT_PRELU, ctx.P('block1.1.weight'), gs, acc)
dxr = engine.conv_bwd(ctx, s_b1, db1pre, need_dx=True)
print('dxr rel', rel_err(nchw(dxr), aux['rectified'].grad), 'norms', float(dxr.norm()), float(aux['rectified'].grad.norm()))
dctrl = ops.tps_sample_bwd(dxr, xn, src, ctx.P('tps.inverse_kernel'), ctx.P('tps.reprT'))
print('dctrl rel', rel_err(dctrl, aux['ctrl'].grad))
# tps bwd with the ORACLE's dxr
dxr_o = aux['rectified'].grad.permute(0, 2, 3, 1).contiguous().cuda()
dctrl2 = ops.tps_sample_bwd(dxr_o, xn, src, ctx.P('tps.inverse_kernel'), ctx.P('tps.reprT'))
print('dctrl (oracle dxr) rel', rel_err(dctrl2, aux['ctrl'].grad))
print('block1 weight grad rel', rel_err(st.gviews['block1.0.weight'], leaves['block1.0.weight'].grad))
# clamp statistics
s = src.cpu()
print('src out of [0,1]: x %.3f y %.3f' % (float(((s[..., 0] < 0) | (s[..., 0] > 1)).float().mean()), float(((s[..., 1] < 0) | (s[..., 1] > 1)).float().mean())))
