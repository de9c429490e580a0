"""GPU-box microbenchmark: CUDA-event timing of the conv kernels — per-tap TMA (conv_tc / wgrad_tc) vs padded-flat (conv_flat)."""
import os, sys
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
only = sys.argv[1] if len(sys.argv) > 1 and sys.argv[1] != 'all' else None
for (B, Ci, Co, H, W, k) in [(48, 64, 64, 16, 64, 3), (48, 64, 64, 16, 64, 1), (48, 64, 256, 16, 64, 3), (48, 64, 192, 16, 64, 1),
                             (48, 192, 64, 16, 64, 1), (48, 64, 4, 32, 128, 9), (192, 64, 64, 16, 64, 3)]:
    x = torch.randn(B, H, W, Ci, device=d); w = torch.randn(Co, Ci, k, k, device=d) * 0.05; b = torch.zeros(Co, device=d)
    dy = torch.randn(B, H, W, Co, device=d)
    flops = 2.0 * B * H * W * Co * Ci * k * k
    pad = k // 2
    for mode in ('bf16x3', 'bf16'):
        if only and mode != only: continue
        ops.CONV_MODE = mode
        lo = mode == 'bf16x3'
        (wf, _) = ops.pack_conv_weights_bf16(w, False, lo)
        y = torch.empty(B, H, W, Co, device=d)
        tag = f'{mode:7s} B={B} {Ci}->{Co} {H}x{W} k{k}'
        if ops.tc_supported(B, H, W, Ci, H, W, Co, k, k):
            planes = ops.split_bf16(x, lo)
            f = lambda: ops.conv_tc(planes, wf, b, B, H, W, Ci, Co, k, k, pad, pad, out=y)
            med, mn = timeit(f, flush=flush); hot, hmn = timeit(f)
            print(f'tc    {tag}: cold {med:7.1f} us  hot {hot:7.1f} us (min {hmn:6.1f})  {flops / hot / 1e6:8.1f} TFLOP/s hot', flush=True)
        if ops.flat_supported(B, H, W, pad, Ci, Co, k, k, pad, pad):
            fp = ops.split_bf16_flat(x, pad, lo)
            f = lambda: ops.conv_flat(fp, wf, b, Co, k, k, pad, pad, out=y)
            med, mn = timeit(f, flush=flush); hot, hmn = timeit(f)
            print(f'flat  {tag}: cold {med:7.1f} us  hot {hot:7.1f} us (min {hmn:6.1f})  {flops / hot / 1e6:8.1f} TFLOP/s hot', flush=True)
        if Ci == 64:
            dw = torch.zeros(Co, Ci, k, k, device=d)
            cpad = (Co + 63) // 64 * 64
            if ops.wgrad_tc_supported(B, H, W, Ci, Co, k, k):
                xp = ops.split_bf16(x, lo)
                dyp = ops.split_bf16(dy, lo) if cpad == Co else ops.split_bf16_pad(dy, cpad, lo)
                f = lambda: ops.conv_wgrad_tc(xp, dyp, dw, B, H, W, Ci, Co, k, k, pad, pad, False)
                hot, hmn = timeit(f)
                print(f'wg-tc   {tag}: hot {hot:7.1f} us (min {hmn:6.1f})  {flops / hot / 1e6:8.1f} TFLOP/s', flush=True)
            if ops.wgrad_flat_supported(B, H, W, pad, Ci, Co, k, k):
                xp = ops.split_bf16_flat(x, pad, lo)
                dyp = ops.split_bf16_flat(dy, pad, lo, cpad)
                f = lambda: ops.conv_wgrad_flat(xp, dyp, dw, Co, k, k, pad, pad, False)
                hot, hmn = timeit(f)
                print(f'wg-flat {tag}: hot {hot:7.1f} us (min {hmn:6.1f})  {flops / hot / 1e6:8.1f} TFLOP/s', flush=True)
    if only is None or only == 'bf16x3':
        f = lambda: ops.split_bf16(x, True)
        hot, _ = timeit(f)
        f2 = lambda: ops.split_bf16_flat(x, pad, True)
        hot2, _ = timeit(f2)
        print(f'split_bf16 {x.numel() * 4 / 1e6:.1f} MB in: dense {hot:6.1f} us, flat(P={pad}) {hot2:6.1f} us', flush=True)
