"""GPU-box microbenchmark: what one dependent kernel node costs inside a replayed CUDA graph (launch-to-launch time of a chain of trivial
kernels), and the same chain with a ~5 us kernel — the floor under an 800-launch train step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from tpgsr_b200 import ops
d = torch.device('cuda:0')
cnt = torch.zeros(4, dtype=torch.int64, device=d)
a = torch.randn(48 * 1024 * 64, device=d); b = torch.randn_like(a); c = torch.empty_like(a)
for name, fn, n in (('add_i64 (1 thread of work)', lambda: ops.add_i64(cnt, 1), 1000), ('add [12.6 MB + 12.6 MB -> 12.6 MB]', lambda: ops.add(a, b, out=c), 400)):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); e1.synchronize()
    print(f'{name}: {e0.elapsed_time(e1) * 1e3 / (5 * n):.2f} us per graph node', flush=True)
