"""GPU-box: cluster LSTM forward / backward per-launch time (CUDA graph), tensor-core kernels against the FFMA cluster kernels on the
same inputs (tpgsr_lstm_set_mode), and the per-step timeline of the forward kernel's first CTA."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from tpgsr_b200 import ops_prior
from tpgsr_b200._lib import LIB
d = torch.device('cuda:0')


def batch_time(f, n=6, reps=5):
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


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


T = 26
for B in (48, 5):
    gx = torch.randn(B * T, 2048, device=d) * 0.5
    whh = torch.randn(2, 1024, 256, device=d) * 0.06
    bhh = torch.randn(2, 1024, device=d) * 0.1
    whhT = ops_prior.lstm_transpose_whh(whh)
    dy = torch.randn(B * T, 512, device=d)
    outs = {}
    for mode in (0, 1):
        LIB.tpgsr_lstm_set_mode(mode)
        f = lambda: ops_prior.lstm_fwd(gx, whhT, bhh, T, B, True, batch_first=True)
        y, stash = f()
        b = lambda: ops_prior.lstm_bwd(dy, stash, whh, T, B, batch_first=True)
        dgx = b()
        torch.cuda.synchronize()
        outs[mode] = (y.clone(), stash.clone(), dgx.clone())
        print(f'B={B} {"tensor-core" if mode else "FFMA       "}: lstm fwd {batch_time(f):6.1f} us, bwd {batch_time(b):6.1f} us per launch (T = {T})', flush=True)
        dbg = torch.zeros(T * 6, dtype=torch.int64, device=d)
        LIB.tpgsr_lstm_set_debug(dbg.data_ptr()); f(); torch.cuda.synchronize(); LIB.tpgsr_lstm_set_debug(None)
        t = dbg.cpu().view(T, 6)
        per = (t[1:, 0] - t[:-1, 0]).float()
        names = ['mat-vec phase done', 'partials exchanged (sync)', 'gates done (sync)', 'h pushed', 'h of the previous step landed (wait at the step start; FFMA kernel: cluster wait at the step end)']
        print(f'   forward timeline, cycles per step median {per.median().item():.0f}: ' +
              ', '.join(f'{n} +{(t[:, i + 1] - t[:, 0]).float().median().item():.0f}' for i, n in enumerate(names)), flush=True)
    print('   tensor-core vs FFMA, rel L2: ' + '  '.join(f'{n} {rel(a, b_):.1e}' for n, a, b_ in zip(('y', 'stash', 'dgx'), outs[1], outs[0])), flush=True)
LIB.tpgsr_lstm_set_mode(1)
