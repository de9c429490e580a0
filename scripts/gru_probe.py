"""GPU-box: which part of the tensor-core GRU step costs what (debug skip bits of tpgsr_gru_set_mode; results are garbage, timing only)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from tpgsr_b200 import ops
from tpgsr_b200._lib import LIB
sys.path.insert(0, os.path.join(ROOT, 'scripts'))
d = torch.device('cuda:0')
def batch_time(f, n=10, reps=5):
    for _ in range(3): f()
    torch.cuda.synchronize()
    st = torch.cuda.Stream(); g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        f(); torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=st):
            for _ in range(n): f()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / n)
    return best
for B in (5, 48):
    H, W = 16, 64
    M = B * H * W
    gx = torch.randn(M, 192, device=d) * 0.3; whh = torch.randn(2, 96, 32, device=d) * 0.1; bhh = torch.zeros(2, 96, device=d); res = torch.randn(M, 64, device=d)
    for vertical in (False, True):
        for name, bits in (('full', 0), ('no stores', 1), ('no fetch', 2), ('no MMA', 4), ('no MUFU gates', 8), ('no cp.async wait', 16), ('no stores+fetch+wait', 19),
                           ('only barrier+ldmatrix+split', 31)):
            LIB.tpgsr_gru_set_mode(1 | (bits << 4))
            f = lambda: ops.gru_fwd(gx, whh, bhh, B, H, W, vertical, residual=res)
            print(f'B={B} {"vertical T=16" if vertical else "horizontal T=64"} {name:28s} {batch_time(f):7.1f} us', flush=True)
LIB.tpgsr_gru_set_mode(1)
