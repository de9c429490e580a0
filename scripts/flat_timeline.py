"""GPU-box: per-CTA clock64 timeline of conv_flat (developer instrumentation) + back-to-back event timing (20 launches per event pair)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from tpgsr_b200 import ops
from tpgsr_b200._lib import LIB
d = torch.device('cuda:0')
names = ['setup_done', 'A_full', 'B0_full', 'mma_issued', 'tmem_full0', 'tile0_stored', 'epi_done', 'item0_ldtm', 'item0_staged']

def batch_time(f, n=20, reps=5):
    for _ in range(3): f()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): f()
        e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / n)
    return best

for (B, Ci, Co, H, W, k) in [(48, 64, 64, 16, 64, 1), (48, 64, 64, 16, 64, 3), (48, 64, 4, 32, 128, 9)]:
    pad = k // 2
    x = torch.randn(B, H, W, Ci, device=d); w = torch.randn(Co, Ci, k, k, device=d) * 0.05; b = torch.zeros(Co, device=d)
    for mode in ('bf16x3', 'bf16'):
        ops.CONV_MODE = mode
        lo = mode == 'bf16x3'
        fp = ops.split_bf16_flat(x, pad, lo); (wf, _) = ops.pack_conv_weights_bf16(w, False, lo)
        y = torch.empty(B, H, W, Co, device=d)
        f = lambda: ops.conv_flat(fp, wf, b, Co, k, k, pad, pad, out=y)
        us = batch_time(f)
        dbg = torch.zeros(4096, 16, dtype=torch.int64, device=d)
        LIB.tpgsr_conv2d_flat_set_debug(dbg.data_ptr())
        f(); torch.cuda.synchronize()
        LIB.tpgsr_conv2d_flat_set_debug(None)
        t = dbg.cpu(); t = t[t[:, 0] != 0]
        rel = (t[:, 1:10] - t[:, 0:1]).float()
        print(f'k={k} {Ci}->{Co} {mode}: {us:.1f} us/launch back-to-back (L2-hot); per-CTA cycles since CTA start, median / p90 over {t.shape[0]} CTAs; '
              f'first CTA start spread {(t[:, 0].max() - t[:, 0].min()).item()} cyc', flush=True)
        for i, n in enumerate(names):
            col = rel[:, i]
            print(f'   {n:12s} {col.median().item():9.0f} {col.kthvalue(max(1, int(0.9 * len(col)))).values.item():9.0f}', flush=True)
        if k == 3 and lo:
            xp = ops.split_bf16(x, lo)
            g = lambda: ops.conv_tc(xp, wf, b, B, H, W, Ci, Co, k, k, pad, pad, out=y)
            print(f'   (per-tap conv_tc, same shape: {batch_time(g):.1f} us/launch back-to-back)')
            dy = torch.randn(B, H, W, Co, device=d); dw = torch.zeros(Co, Ci, k, k, device=d)
            dyp = ops.split_bf16_flat(dy, pad, lo)
            h = lambda: ops.conv_wgrad_flat(fp, dyp, dw, Co, k, k, pad, pad, False)
            print(f'   (wgrad_flat + reduce, same shape: {batch_time(h):.1f} us/launch pair back-to-back)')
            s_ = lambda: ops.split_bf16_flat(x, pad, lo)
            print(f'   (split_bf16_flat: {batch_time(s_):.1f} us/launch back-to-back)')
