"""GPU-box: weights-stationary conv plan (conv_ws_kernel) vs the flat plan on the 3x3 64->64 bench shape: correctness, back-to-back and
L2-flushed event timing, per-CTA clock64 timeline (developer instrumentation)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from tpgsr_b200 import ops
from tpgsr_b200._lib import LIB
d = torch.device('cuda:0')
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
names = (['setup_done', 'box0_landed', 'w_in_tmem'] + [f'mma_issued_t{t}' for t in range(4)] + [f'acc_full_t{t}' for t in range(4)] +
         [f'ldtm_done_t{t}' for t in range(4)] + [f'staged_t{t}' for t in range(4)] + [f'stored_t{t}' for t in range(4)] + ['cta_done'])


def batch_time(f, n=20, reps=5):
    """device time per launch: n launches captured in one CUDA graph (no host launch cost in the figure)"""
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


def flushed_time(f, reps=20):
    junk = torch.empty(256 << 20, dtype=torch.uint8, device=d)
    ts = []
    for _ in range(reps):
        junk.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


K = int(os.environ.get('WS_K', '3'))
for (B, H, W) in [(48, 16, 64), (8, 16, 64), (192, 16, 64)]:
    x = torch.randn(B, H, W, 64, device=d); w = torch.randn(64, 64, K, K, device=d) / (8.0 * K); b = torch.randn(64, device=d) * 0.1
    want = F.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), b.double(), padding=K // 2).permute(0, 2, 3, 1).float()
    fp = ops.split_bf16_flat(x, K // 2, True); (wf, _) = ops.pack_conv_weights_bf16(w, False, True)
    y = torch.empty(B, H, W, 64, device=d)
    f = lambda: ops.conv_flat(fp, wf, b, 64, K, K, K // 2, K // 2, out=y)
    for tag, bits in (('weights-stationary', 0), ('flat plan', 16)):
        LIB.tpgsr_conv2d_flat_set_debug_skip(bits)
        y.zero_(); f(); torch.cuda.synchronize()
        err = float((y - want).norm() / want.norm())
        print(f'B={B} {tag}: rel err vs fp64 conv {err:.2e}; {batch_time(f):.1f} us back-to-back in a CUDA graph (L2-hot), {flushed_time(f):.1f} us L2-flushed', flush=True)
        if bits != 16:
            dbg = torch.zeros(4096, 32, dtype=torch.int64, device=d)
            LIB.tpgsr_conv2d_flat_set_debug(dbg.data_ptr())
            f(); torch.cuda.synchronize()
            LIB.tpgsr_conv2d_flat_set_debug(None)
            t = dbg.cpu(); t = t[t[:, 0] != 0]
            rel = (t[:, 1:25] - t[:, 0:1]).float()
            print(f'   per-CTA cycles since CTA start, median / p90 over {t.shape[0]} CTAs; CTA start spread {(t[:, 0].max() - t[:, 0].min()).item()} cyc')
            for i, n in enumerate(names):
                col = rel[:, i]
                col = col[col > 0]
                if len(col): print(f'   {n:14s} {col.median().item():9.0f} {col.kthvalue(max(1, int(0.9 * len(col)))).values.item():9.0f}', flush=True)
    LIB.tpgsr_conv2d_flat_set_debug_skip(0)
