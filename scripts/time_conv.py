"""GPU-box microbenchmark: CUDA-event timing of the conv kernels (tensor-core and CUDA-core paths)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tpgsr_b200 import ops
d = torch.device('cuda:0')

def timeit(fn, n=20, flush=None):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        if flush is not None: flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]

flush = torch.empty(256 << 20, dtype=torch.uint8, device=d)
only = sys.argv[1] if len(sys.argv) > 1 else None
for (B, Ci, Co, H, W, k) in [(48, 64, 64, 16, 64, 3), (48, 64, 64, 16, 64, 1), (48, 64, 256, 16, 64, 3), (48, 64, 192, 16, 64, 1), (48, 64, 4, 32, 128, 9), (192, 64, 64, 16, 64, 3)]:
    x = torch.randn(B, H, W, Ci, device=d); w = torch.randn(Co, Ci, k, k, device=d) * 0.05; b = torch.zeros(Co, device=d)
    flops = 2.0 * B * H * W * Co * Ci * k * k
    for mode in ('bf16x3', 'bf16'):
        if only and mode != only: continue
        ops.CONV_MODE = mode
        planes = ops.split_bf16(x, mode == 'bf16x3')
        (wf, _) = ops.pack_conv_weights_bf16(w, False, mode == 'bf16x3')
        y = torch.empty(B, H, W, Co, device=d)
        f = lambda: ops.conv_tc(planes, wf, b, B, H, W, Ci, Co, k, k, k // 2, k // 2, out=y)
        med, mn = timeit(f, flush=flush); hot, hmn = timeit(f)
        print(f'tc   {mode:7s} B={B} {Ci}->{Co} {H}x{W} k{k}: cold {med:7.1f} us  hot {hot:7.1f} us (min {hmn:6.1f})  {flops / hot / 1e6:8.1f} TFLOP/s hot')
    wf32, _ = ops.pack_conv_weights(w, False)
    y = torch.empty(B, H, W, Co, device=d)
    f = lambda: ops.conv_fwd(x, wf32, b, Co, k, k, k // 2, k // 2, out=y)
    hot, hmn = timeit(f)
    print(f'simt fp32    B={B} {Ci}->{Co} {H}x{W} k{k}: hot {hot:7.1f} us  {flops / hot / 1e6:8.1f} TFLOP/s')
    f = lambda: ops.split_bf16(x, True)
    hot, _ = timeit(f)
    print(f'split_bf16   {x.numel() * 4 / 1e6:.1f} MB in: hot {hot:6.1f} us')
