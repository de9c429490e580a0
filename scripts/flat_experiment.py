"""GPU-box experiment: where does the conv_flat epilogue time go?  Times the 3x3 / 1x1 kernels with parts of the epilogue disabled."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from tpgsr_b200 import ops
from tpgsr_b200._lib import LIB
d = torch.device('cuda:0')
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
B, H, W = 48, 16, 64
for k in (1, 3):
    pad = k // 2
    x = torch.randn(B, H, W, 64, device=d); w = torch.randn(64, 64, k, k, device=d) * 0.05; b = torch.zeros(64, device=d)
    fp = ops.split_bf16_flat(x, pad, True); (wf, _) = ops.pack_conv_weights_bf16(w, False, True)
    y = torch.empty(B, H, W, 64, device=d)
    f = lambda: ops.conv_flat(fp, wf, b, 64, k, k, pad, pad, out=y)
    for bits, name in ((0, 'full'), (1, 'no global stores'), (2, 'no TMEM load / staging / stores')):
        LIB.tpgsr_conv2d_flat_set_debug_skip(bits)
        print(f'k={k} {name:32s}: {batch_time(f):6.1f} us', flush=True)
    LIB.tpgsr_conv2d_flat_set_debug_skip(0)
# empty-kernel launch floor for reference
z = torch.zeros(1, device=d)
print(f'launch floor (tiny add_i64 kernel): {batch_time(lambda: ops.add_i64(torch.zeros(1, dtype=torch.int64, device=d), 1)):6.1f} us')
