"""GPU-box microbenchmark: GRU recurrence kernels (forward; backward recurrence; dW_hh + reduce) at the bench shape, tensor-core kernels
against the warp-per-sequence FFMA kernels of round 1 (tpgsr_gru_set_mode), device time per launch in a CUDA graph + outputs compared."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from tpgsr_b200 import ops
from tpgsr_b200._lib import LIB
d = torch.device('cuda:0')


def batch_time(f, n=10, reps=5):
    for _ in range(3): f()
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        f()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=st):
            for _ in range(n): f()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / n)
    return best


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


for B in (48, 5):
    H, W = 16, 64
    M = B * H * W
    g_ = torch.Generator(device='cpu').manual_seed(1)
    gx = torch.randn(M, 192, generator=g_).to(d); whh = (torch.randn(2, 96, 32, generator=g_) * 0.18).to(d); bhh = (torch.randn(2, 96, generator=g_) * 0.1).to(d)
    res = torch.randn(M, 64, generator=g_).to(d); dy = torch.randn(M, 64, generator=g_).to(d)
    for vertical in (True, False):
        outs = {}
        for mode in (0, 2):
            LIB.tpgsr_gru_set_mode(mode)
            f = lambda: ops.gru_fwd(gx, whh, bhh, B, H, W, vertical, residual=res)
            y, ysum, stash = f()
            t_f = batch_time(f)
            r = lambda: ops.gru_bwd_recurrence(dy, y, stash, whh, B, H, W, vertical)
            dgx, part = r()
            t_b = batch_time(r)
            dwhh, dbhh, dbih = torch.zeros_like(whh), torch.zeros_like(bhh), torch.zeros(192, device=d)
            w = lambda: ops.gru_bwd_weights(dgx, y, stash, part, dwhh, dbhh, dbih, False, B, H, W, vertical)
            w()
            t_w = batch_time(w)
            torch.cuda.synchronize()
            outs[mode] = (y.clone(), ysum.clone(), stash.clone(), dgx.clone(), dwhh.clone(), dbhh.clone(), dbih.clone())
            print(f'B={B} GRU {"vertical T=16" if vertical else "horizontal T=64"} {"tensor-core" if mode else "FFMA       "}: fwd {t_f:6.1f} us   '
                  f'bwd recurrence {t_b:6.1f} us   dW_hh + reduce {t_w:6.1f} us', flush=True)
        names = ('y', 'ysum', 'stash', 'dgx', 'dwhh', 'dbhh', 'dbih')
        print('   tensor-core vs FFMA, rel L2: ' + '  '.join(f'{n} {rel(a, b):.1e}' for n, a, b in zip(names, outs[2], outs[0])), flush=True)
LIB.tpgsr_gru_set_mode(2)
# per-step timeline of the tensor-core forward kernel (first CTA, compute warp 0)
B, H, W = 48, 16, 64
M = B * H * W
gx = torch.randn(M, 192, device=d) * 0.3; whh = torch.randn(2, 96, 32, device=d) * 0.1; bhh = torch.zeros(2, 96, device=d); res = torch.randn(M, 64, device=d)
for vertical in (False, True):
    dbg = torch.zeros(64 * 4, dtype=torch.int64, device=d)
    ops.gru_fwd(gx, whh, bhh, B, H, W, vertical, residual=res); torch.cuda.synchronize()
    LIB.tpgsr_gru_set_debug(dbg.data_ptr())
    ops.gru_fwd(gx, whh, bhh, B, H, W, vertical, residual=res); torch.cuda.synchronize()
    LIB.tpgsr_gru_set_debug(None)
    T = 16 if vertical else 64
    t = dbg.cpu().view(64, 4)[:T]
    per = (t[1:, 0] - t[:-1, 0]).float()
    print(f'timeline {"vertical" if vertical else "horizontal"}: cycles per step median {per.median().item():.0f} (min {per.min().item():.0f}, max {per.max().item():.0f}); '
          f'within a step: MMAs done +{(t[:, 1] - t[:, 0]).float().median().item():.0f}, gates done +{(t[:, 2] - t[:, 0]).float().median().item():.0f}, '
          f'stores issued +{(t[:, 3] - t[:, 0]).float().median().item():.0f}', flush=True)
LIB.tpgsr_gru_set_mode(1)
