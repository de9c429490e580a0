"""GPU-box: per-CTA clock64 timeline of conv_tc (developer instrumentation)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from tpgsr_b200 import ops
from tpgsr_b200._lib import LIB
d = torch.device('cuda:0')
for (B, Ci, Co, H, W, k) in [(48, 64, 64, 16, 64, 1), (48, 64, 64, 16, 64, 3)]:
    x = torch.randn(B, H, W, Ci, device=d); w = torch.randn(Co, Ci, k, k, device=d) * 0.05; b = torch.zeros(Co, device=d)
    planes = ops.split_bf16(x, True); (wf, _) = ops.pack_conv_weights_bf16(w, False, True)
    y = torch.empty(B, H, W, Co, device=d)
    for _ in range(3): ops.conv_tc(planes, wf, b, B, H, W, Ci, Co, k, k, k // 2, k // 2, out=y)
    dbg = torch.zeros(B * H * W // 128, 8, dtype=torch.int64, device=d)
    LIB.tpgsr_conv2d_tc_set_debug(dbg.data_ptr())
    ops.conv_tc(planes, wf, b, B, H, W, Ci, Co, k, k, k // 2, k // 2, out=y)
    torch.cuda.synchronize()
    LIB.tpgsr_conv2d_tc_set_debug(None)
    t = dbg.cpu()
    rel = (t[:, 1:7] - t[:, 0:1]).float()
    names = ['setup_done', 'first_full', 'mma_issued', 'tmem_full', 'staged', 'end']
    print(f'k={k}: per-CTA cycles since CTA start (median / p90 over {t.shape[0]} CTAs)')
    for i, n in enumerate(names):
        col = rel[:, i]
        print(f'   {n:12s} {col.median().item():9.0f} {col.kthvalue(int(0.9 * len(col))).values.item():9.0f}')
