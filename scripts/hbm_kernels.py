"""GPU-box microbenchmark of the HBM-bound kernels north_star names (grid_sample / TPS, PixelShuffle, gradient loss, BiLSTM step
recurrence, bicubic+gray), each at the bench batch (B=48: L2-resident tensors) and at a batch whose working set leaves the 126 MB L2
(>= 256 MB), CUDA-event timed with an L2 flush between launches.  Prints achieved GB/s on the ALGORITHMIC bytes of SURVEY §8d next to the
measured HBM peak.  Run the same script under `ncu --set full -k regex:...` for the per-launch DRAM bytes (profiles/r02/README.md).

    python scripts/hbm_kernels.py [--json out.json] [--once]      # --once: one launch per kernel and size (for ncu captures)
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tpgsr_b200 import ops, ops_prior  # noqa: E402

d = torch.device('cuda:0')
ap = argparse.ArgumentParser()
ap.add_argument('--json', default='')
ap.add_argument('--once', action='store_true')
args = ap.parse_args()
peak = 6551.4
pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
if os.path.exists(pk):
    peak = json.load(open(pk))['hbm_gbs']
flush = torch.empty(512 << 20, dtype=torch.uint8, device=d)


def timed(f, n=10):
    if args.once:
        f()
        torch.cuda.synchronize()
        return 0.0
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        f()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


rows = []


def report(name, B, bytes_per_img, us, note=''):
    gbs = B * bytes_per_img / us / 1e3 if us else 0.0
    rows.append({'kernel': name, 'batch': B, 'algorithmic_bytes': B * bytes_per_img, 'us': us, 'GBs': gbs, 'frac_of_hbm_peak': gbs / peak, 'note': note})
    print(f'{name:28s} B={B:6d}  {B * bytes_per_img / 1e6:9.1f} MB  {us:9.1f} us  {gbs:8.0f} GB/s  = {gbs / peak:5.2f} of {peak:.0f}  {note}', flush=True)


from oracle.ops import tps_buffers  # noqa: E402  (test infrastructure: only the constant TPS buffers, to have valid operands)

buf = tps_buffers(16, 64)
inv = buf['inverse_kernel'].to(d)
reprT = buf['target_coordinate_repr'].t().contiguous().to(d)
tcp = buf['target_control_points'].to(d)


def _section0(B):
    # ---- TPS + grid_sample (model/tps_spatial_transformer.py:97-112): 40 960 B/img fwd, 49 152 B/img bwd
    img = torch.rand(B, 16, 64, 4, device=d)
    ctrl = (tcp[None] + 0.02 * torch.randn(B, 20, 2, device=d)).contiguous()
    out, src = ops.tps_sample_fwd(img, ctrl, inv, reprT)
    report('tps_sample_fwd', B, 40960, timed(lambda: ops.tps_sample_fwd(img, ctrl, inv, reprT)))
    dout = torch.randn_like(out)
    report('tps_sample_bwd', B, 49152, timed(lambda: ops.tps_sample_bwd(dout, img, src, inv, reprT)))
    del img, out, src, dout


for B in (48, 8192):
    try:
        _section0(B)
    except Exception as e:
        print('section 0 failed at B =', B, repr(e), flush=True)
        torch.cuda.synchronize()


def _section1(B):
    # ---- PixelShuffle + mish (model/tsrn.py:464-477): 2 097 152 B/img (1 MiB in + 1 MiB out); bwd reads dy + x, writes dx: 3 MiB
    u = torch.randn(B, 16, 64, 256, device=d)
    v = ops.pixelshuffle_mish_fwd(u)
    report('pixelshuffle_mish_fwd', B, 2097152, timed(lambda: ops.pixelshuffle_mish_fwd(u)))
    dv = torch.randn_like(v)
    report('pixelshuffle_mish_bwd', B, 3145728, timed(lambda: ops.pixelshuffle_mish_bwd(dv, u)))
    del u, v, dv


for B in (48, 192):
    try:
        _section1(B)
    except Exception as e:
        print('section 1 failed at B =', B, repr(e), flush=True)
        torch.cuda.synchronize()


def _section2(B):
    # ---- ImageLoss = MSE + 1e-4 gradient-map L1 (loss/image_loss.py:19-51): 131 072 B/img fwd, bwd +65 536 (reads both, writes grad)
    sr = torch.rand(B, 4, 32, 128, device=d)
    hr = torch.rand(B, 4, 32, 128, device=d)
    loss = torch.zeros(1, device=d)
    ops.workspace(256 << 20, d)
    report('image_loss_fwd', B, 131072, timed(lambda: ops.image_loss_fwd(sr, hr, True, 1.0, 1e-4, 100.0, loss=loss)), '2 kernels (partials + finalize)')
    report('image_loss_bwd', B, 196608, timed(lambda: ops.image_loss_bwd(sr, hr, True, 1.0, 1e-4, 100.0)))
    # ---- parse_crnn_data from the HR/SR map (interfaces/base.py:806-829): 61 952 B/img; backward reads dgray 12 800, writes 3x16 384
    report('bicubic_gray_fwd (32x128)', B, 61952, timed(lambda: ops_prior.bicubic_gray_fwd(sr)))
    g = ops_prior.bicubic_gray_fwd(sr)
    dsr = torch.zeros_like(sr)
    report('bicubic_gray_bwd (32x128)', B, 12800 + 49152, timed(lambda: ops_prior.bicubic_gray_bwd(g, dsr, False)))
    del sr, hr, g, dsr


for B in (48, 4096):
    try:
        _section2(B)
    except Exception as e:
        print('section 2 failed at B =', B, repr(e), flush=True)
        torch.cuda.synchronize()


def _section3(B):
    # ---- BiLSTM(256) step recurrence (model/crnn/crnn.py:5-26), T = 26: reads gx [T*B,2048], writes y [T*B,512] (+ stash [T*B,2,5,256]);
    #      serial over T: latency-bound, bytes/s reported for the record
    T = 26
    gx = torch.randn(T * B, 2048, device=d) * 0.5
    whh = torch.randn(2, 1024, 256, device=d) * 0.05
    whhT = ops_prior.lstm_transpose_whh(whh)
    bhh = torch.zeros(2, 1024, device=d)
    y, stash = ops_prior.lstm_fwd(gx, whhT, bhh, T, B, True, batch_first=True)
    per_img_f = T * (2048 + 512 + 2560) * 4
    report('lstm_fwd (T=26,H=256,bi)', B, per_img_f, timed(lambda: ops_prior.lstm_fwd(gx, whhT, bhh, T, B, True, batch_first=True)), 'incl. stash write')
    dy = torch.randn(T * B, 512, device=d)
    per_img_b = T * (512 + 2560 + 2048) * 4
    report('lstm_bwd (T=26,H=256,bi)', B, per_img_b, timed(lambda: ops_prior.lstm_bwd(dy, stash, whh, T, B, batch_first=True)))
    del gx, y, stash, dy


for B in (48, 1024):
    try:
        _section3(B)
    except Exception as e:
        print('section 3 failed at B =', B, repr(e), flush=True)
        torch.cuda.synchronize()


def _section4(B):
    # ---- GruBlock recurrence (model/tsrn.py:491-508): SURVEY counts 524 288 B/img for the whole block; the recurrence kernel itself moves
    #      gx 786 432 + y 262 144 (+ residual, ysum, stash 1 048 576) per image
    H, W = 16, 64
    M = B * H * W
    gx = torch.randn(M, 192, device=d)
    whh = torch.randn(2, 96, 32, device=d) * 0.1
    bhh = torch.zeros(2, 96, device=d)
    res = torch.randn(M, 64, device=d)
    dy = torch.randn(M, 64, device=d)
    for vertical in (True, False):
        tag = 'v' if vertical else 'h'
        y, ysum, stash = ops.gru_fwd(gx, whh, bhh, B, H, W, vertical, residual=res)
        report(f'gru_fwd ({tag})', B, 524288, timed(lambda: ops.gru_fwd(gx, whh, bhh, B, H, W, vertical, residual=res)), 'vs the GruBlock algorithmic bytes')
        report(f'gru_bwd_recurrence ({tag})', B, 524288, timed(lambda: ops.gru_bwd_recurrence(dy, y, stash, whh, B, H, W, vertical)), 'vs the GruBlock algorithmic bytes')
    del gx, res, dy, y, ysum, stash


for B in (48, 384):
    try:
        _section4(B)
    except Exception as e:
        print('section 4 failed at B =', B, repr(e), flush=True)
        torch.cuda.synchronize()


def _section5(B):
    # ---- CRNN pool0 (model/crnn/crnn.py:40: MaxPool2d(2, 2) on [B,32,100,64]): reads 819 200 B/img, writes 204 800 + 204 800 (argmax) B/img;
    #      backward reads dy + argmax, writes dx
    x = torch.randn(B, 32, 100, 64, device=d)
    p, arg = ops.maxpool_fwd(x, (2, 2), (2, 2), (0, 0), need_argmax=True)
    report('maxpool_fwd (pool0)', B, 819200 + 2 * 204800, timed(lambda: ops.maxpool_fwd(x, (2, 2), (2, 2), (0, 0), need_argmax=True)))
    dp = torch.randn_like(p)
    report('maxpool_bwd (pool0)', B, 819200 + 2 * 204800, timed(lambda: ops.maxpool_bwd(dp, arg, tuple(x.shape), (2, 2), (2, 2), (0, 0))))
    # ---- CRNN conv0 (model/crnn/crnn.py:35: 1 -> 64, 3x3, ReLU): reads 12 800 B/img, writes 819 200 B/img
    g = torch.randn(B, 32, 100, 1, device=d)
    w = torch.randn(64, 1, 3, 3, device=d) * 0.3
    b = torch.zeros(64, device=d)
    wf, _ = ops.pack_conv_weights(w, False)
    report('conv0 1->64 3x3 + ReLU', B, 12800 + 819200, timed(lambda: ops.conv_fwd(g, wf, b, 64, 3, 3, 1, 1, act=ops.ACT_RELU)))
    del x, p, arg, dp, g


for B in (48, 384):
    try:
        _section5(B)
    except Exception as e:
        print('section 5 failed at B =', B, repr(e), flush=True)
        torch.cuda.synchronize()


def _section6():
    # ---- clip + Adam over the config-3 arena (11.9 M parameters): 28 B/param
    from tpgsr_b200 import optim as topt
    n = 11_900_000
    flat = torch.randn(n, device=d) * 0.01
    gflat = torch.randn(n, device=d) * 0.001
    opt = topt.FlatAdam(flat, n_clipped=n // 4, grads=gflat)
    us = timed(lambda: opt.step())
    rows.append({'kernel': 'clip + adam (3 launches)', 'batch': 0, 'algorithmic_bytes': n * 28, 'us': us, 'GBs': n * 28 / us / 1e3 if us else 0.0})
    print(f'clip + adam, 11.9 M params     {n * 28 / 1e6:9.1f} MB  {us:9.1f} us  {(n * 28 / us / 1e3 if us else 0):8.0f} GB/s', flush=True)


try:
    _section6()
except Exception as e:
    print('section 6 failed', repr(e), flush=True)
    torch.cuda.synchronize()
if args.json:
    json.dump({'hbm_peak_gbs': peak, 'rows': rows}, open(args.json, 'w'), indent=1)
