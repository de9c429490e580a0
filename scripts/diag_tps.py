"""Diagnostic (GPU box): TPS backward kernel with CPU-computed src vs golden."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from conftest import load_golden, rel_err
from tpgsr_b200 import ops
g = load_golden('tps')
d = torch.device('cuda:0')
nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous()
inv = torch.from_numpy(g['tps.inverse_kernel']).to(d)
reprT = torch.from_numpy(g['tps.target_coordinate_repr']).t().contiguous().to(d)
img = nhwc(torch.from_numpy(g['img'])).to(d)
w = nhwc(torch.from_numpy(g['w'])).to(d)
ctrl = torch.from_numpy(g['ctrl']).to(d)
out, src = ops.tps_sample_fwd(img, ctrl, inv, reprT)
src_ref = torch.from_numpy(g['src']).to(d)
print('src max abs diff', float((src - src_ref).abs().max()))
for name, s in (('gpu src', src), ('ref src', src_ref)):
    dc = ops.tps_sample_bwd(w, img, s.contiguous(), inv, reprT)
    print(name, 'rel', rel_err(dc, g['grad_ctrl']), dc[0, :3].tolist())
print('ref', g['grad_ctrl'][0, :3].tolist())
# per-batch errors
dc = ops.tps_sample_bwd(w, img, src_ref.contiguous(), inv, reprT).cpu()
ref = torch.from_numpy(g['grad_ctrl'])
print('per-coordinate rel err x/y:', float((dc[..., 0] - ref[..., 0]).norm() / ref[..., 0].norm()), float((dc[..., 1] - ref[..., 1]).norm() / ref[..., 1].norm()))
