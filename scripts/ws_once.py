"""GPU-box: a few launches of the 3x3 64->64 bench conv (B = 48) for an `ncu --set full` capture of the weights-stationary kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from tpgsr_b200 import ops
d = torch.device('cuda:0')
B, H, W = 48, 16, 64
x = torch.randn(B, H, W, 64, device=d); w = torch.randn(64, 64, 3, 3, device=d) / 24.0; b = torch.zeros(64, device=d)
fp = ops.split_bf16_flat(x, 1, True); (wf, _) = ops.pack_conv_weights_bf16(w, False, True)
y = torch.empty(B, H, W, 64, device=d)
for _ in range(4):
    ops.conv_flat(fp, wf, b, 64, 3, 3, 1, 1, out=y)
torch.cuda.synchronize()
print('ok', float(y.abs().mean()))
