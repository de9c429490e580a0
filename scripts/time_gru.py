"""GPU-box microbenchmark: GRU recurrence kernels (forward, backward = BPTT + dW_hh + reduce), 20 launches per CUDA-event pair."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from tpgsr_b200 import ops
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
M = B * H * W
gx = torch.randn(M, 192, device=d); whh = torch.randn(2, 96, 32, device=d) * 0.1; bhh = torch.zeros(2, 96, device=d)
res = torch.randn(M, 64, device=d); dy = torch.randn(M, 64, device=d)
for vertical in (True, False):
    f = lambda: ops.gru_fwd(gx, whh, bhh, B, H, W, vertical, residual=res)
    y, ysum, stash = f()
    t_f = batch_time(f)
    dwhh, dbhh, dbih = torch.zeros_like(whh), torch.zeros_like(bhh), torch.zeros(192, device=d)
    g = lambda: ops.gru_bwd(dy, y, stash, whh, dwhh, dbhh, dbih, False, B, H, W, vertical)
    t_b = batch_time(g)
    fwd_bytes = M * (192 + 64 + 64 + 64 + 256) * 4
    bwd_bytes = M * (64 + 64 + 256 + 192) * 4 + M * (192 + 64 + 64) * 4
    print(f'GRU {"vertical T=16 S=3072" if vertical else "horizontal T=64 S=768"}: fwd {t_f:6.1f} us ({fwd_bytes / t_f / 1e3:6.0f} GB/s algorithmic)   '
          f'bwd (3 kernels) {t_b:6.1f} us ({bwd_bytes / t_b / 1e3:6.0f} GB/s)', flush=True)
