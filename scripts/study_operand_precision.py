"""CPU study (no GPU needed): how much accuracy does each operand format of the tensor-core convolutions cost on the TSRN train step?

Emulates the formats inside the ORACLE's graph (oracle/models.py) by wrapping F.conv2d for the convolutions that run on the tensor cores
(64k input channels, and the two kw-stacked 9x9 convs): operands are rounded exactly as the kernels' split would round them, products are
accumulated in fp32 like the TMEM accumulators.  Backward: dX from exact dY (the backward pass keeps bf16 hi/lo in every candidate), dW from
the SAME rounded activation planes the forward used.  Prints the error of the SR output, the loss and the gradients against plain fp32.

    python scripts/study_operand_precision.py            # ~1 min on 8 threads

Formats:  fp32    reference
          bf16x3  x = hi+lo, w = hi+lo (bf16), hi*hi + hi*lo + lo*hi                 (default parity mode of the kernels)
          bf16    single bf16 plane each                                              (labelled throughput mode)
          f16a    activations ONE fp16 plane, weights bf16 hi+lo: x16*(w_hi + w_lo)   (round-2 candidate: one stacked MMA per K step)
          bf16a   activations ONE bf16 plane, weights bf16 hi+lo
          bwd2    forward as bf16x3; backward with dY as ONE bf16 plane against split weights / activations (2 products per gradient GEMM)
          bwd2w1  as bwd2, and the weight gradient with ONE product (X_hi dY_hi)
          bwd1    forward as bf16x3; the whole backward pass in plain bf16 (one product per gradient GEMM)
"""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

_conv = F.conv2d


def _split(t, dt=torch.bfloat16):
    hi = t.to(dt).float()
    return hi, (t - hi).to(dt).float()


class QConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, padding, mode):
        if mode in ('bf16x3', 'bwd2', 'bwd2w1', 'bwd1'):
            xh, xl = _split(x); wh, wl = _split(w)
            y = _conv(xh, wh, None, padding=padding) + _conv(xh, wl, None, padding=padding) + _conv(xl, wh, None, padding=padding)
            xs = xh + xl
        elif mode == 'bf16':
            xs = x.bfloat16().float()
            y = _conv(xs, w.bfloat16().float(), None, padding=padding)
        elif mode in ('f16a', 'bf16a'):
            xs = x.half().float() if mode == 'f16a' else x.bfloat16().float()
            wh, wl = _split(w)
            y = _conv(xs, wh + wl, None, padding=padding)
        else:
            raise ValueError(mode)
        if b is not None:
            y = y + b.view(1, -1, 1, 1)
        ctx.save_for_backward(xs, w)
        ctx.padding, ctx.has_b, ctx.mode = padding, b is not None, mode
        return y

    @staticmethod
    def backward(ctx, dy):
        xs, w = ctx.saved_tensors
        dyq = dy.bfloat16().float() if ctx.mode in ('bwd2', 'bwd2w1', 'bwd1') else dy   # dY as ONE bf16 plane in both gradient GEMMs
        wq = w.bfloat16().float() if ctx.mode == 'bwd1' else w                    # 'bwd1': the data gradient with ONE product as well
        dx = torch.nn.grad.conv2d_input(xs.shape, wq, dyq, padding=ctx.padding)
        xw = xs.bfloat16().float() if ctx.mode in ('bwd2w1', 'bwd1') else xs      # the weight gradient with ONE product (X_hi dY_hi)
        dw = torch.nn.grad.conv2d_weight(xw, w.shape, dyq, padding=ctx.padding)
        return dx, dw, (dy.sum((0, 2, 3)) if ctx.has_b else None), None, None


def make_conv(mode):
    def conv2d(x, w, b=None, stride=1, padding=0, *a, **k):
        on_tc = (w.shape[1] % 64 == 0) or (w.shape[2] == 9 and min(w.shape[0], w.shape[1]) <= 4)
        if mode == 'fp32' or not on_tc or stride != 1 or a or k:
            return _conv(x, w, b, stride, padding, *a, **k)
        return QConv.apply(x, w, b, padding, mode)
    return conv2d


def main():
    torch.set_num_threads(8)
    from oracle import models, step
    from oracle.ops import tps_buffers
    from oracle.weights import synth_images, synth_state_dict
    shapes = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'shapes.json')))['tsrn_train']
    lr, hr = synth_images(8, seed=50)
    res = {}
    for mode in ('fp32', 'bf16x3', 'bwd2', 'bwd2w1', 'bwd1', 'f16a', 'bf16a', 'bf16'):
        p = synth_state_dict(shapes, 1)
        p.update({'tps.' + k: v for k, v in tps_buffers(16, 64).items()})
        models.F.conv2d = make_conv(mode)
        try:
            res[mode] = step.tsrn_train_step(p, lr, hr, None)
        finally:
            models.F.conv2d = _conv
    ref = res['fp32']
    print(f"{'mode':8s} {'sr rel-L2':>10s} {'sr max/max':>11s} {'loss rel':>10s} {'grad rel-L2 (all)':>18s} {'worst tensor rel-L2 (non-STN)':>30s}")
    for mode, out in res.items():
        if mode == 'fp32':
            continue
        sr_l2 = float((out['sr'] - ref['sr']).norm() / ref['sr'].norm())
        sr_mx = float((out['sr'] - ref['sr']).abs().max() / ref['sr'].abs().max())
        loss = float((out['loss'] - ref['loss']).abs() / ref['loss'].abs())
        num = sum(float((out['grads'][k] - ref['grads'][k]).norm() ** 2) for k in ref['grads'])
        den = sum(float(ref['grads'][k].norm() ** 2) for k in ref['grads'])
        # (conv biases in front of a train-mode BN have an exactly-zero true gradient: pure rounding noise, skipped)
        big = [k for k in ref['grads'] if 'stn_head' not in k and float(ref['grads'][k].norm()) > 1e-4 * float(ref['grads'][k].numel()) ** 0.5 * 1e-3]
        worst = max(((float((out['grads'][k] - ref['grads'][k]).norm() / (ref['grads'][k].norm() + 1e-30)), k) for k in big
                     if not (k.endswith('conv1.bias') or k.endswith('conv2.bias') or k == 'block7.0.bias') or 'gru' in k), default=(0, ''))
        nostn = [k for k in ref['grads'] if 'stn_head' not in k]
        num2 = sum(float((out['grads'][k] - ref['grads'][k]).norm() ** 2) for k in nostn)
        den2 = sum(float(ref['grads'][k].norm() ** 2) for k in nostn)
        worst = (worst[0], worst[1] + f'   | non-STN grads rel-L2 {(num2 / den2) ** 0.5:.2e}')
        print(f'{mode:8s} {sr_l2:10.2e} {sr_mx:11.2e} {loss:10.2e} {(num / den) ** 0.5:18.2e} {worst[0]:12.2e}  {worst[1]}')


if __name__ == '__main__':
    main()
