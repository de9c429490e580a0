"""Oracle op library: explicit CPU fp32 restatements of every op on the TPGSR hot path.

TEST INFRASTRUCTURE — see ``oracle/__init__.py`` for who may import this.

Every function cites the reference call site (paths relative to ``/root/reference``) whose
behaviour under torch 2.11 it restates.  Backward passes are obtained with autograd over
these explicit forwards (the forwards contain no fused library modules, so the backward
is the textbook chain rule of the written-out math).
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# ----------------------------------------------------------------------------------------
# elementwise
# ----------------------------------------------------------------------------------------

def softplus(x: Tensor) -> Tensor:
    """F.softplus(beta=1, threshold=20) as used by mish (model/tsrn.py:485-488)."""
    return torch.where(x > 20.0, x, torch.log1p(torch.exp(torch.clamp(x, max=20.0))))


def mish(x: Tensor) -> Tensor:
    """x * tanh(softplus(x)) — model/tsrn.py:480-488."""
    return x * torch.tanh(softplus(x))


def prelu(x: Tensor, slope: Tensor) -> Tensor:
    """nn.PReLU() with one shared slope — model/tsrn.py:28."""
    return torch.where(x >= 0, x, slope.reshape(1, -1, 1, 1) * x) if slope.numel() > 1 \
        else torch.where(x >= 0, x, slope.reshape(()) * x)


# ----------------------------------------------------------------------------------------
# normalisation
# ----------------------------------------------------------------------------------------

def batchnorm(x: Tensor, weight: Tensor, bias: Tensor, running_mean: Tensor, running_var: Tensor,
              training: bool, momentum: float = 0.1, eps: float = 1e-5,
              update_running: bool = True) -> Tensor:
    """nn.BatchNorm2d / BatchNorm1d semantics: training normalises with the
    biased batch variance and updates running stats with the unbiased one, in place."""
    dims = [0] + list(range(2, x.dim()))
    shape = [1, -1] + [1] * (x.dim() - 2)
    if training:
        mean = x.mean(dims)
        var = ((x - mean.reshape(shape)) ** 2).mean(dims)
        if update_running:
            n = x.numel() // x.shape[1]
            with torch.no_grad():
                running_mean.mul_(1 - momentum).add_(momentum * mean.detach())
                running_var.mul_(1 - momentum).add_(momentum * var.detach() * (n / max(n - 1, 1)))
    else:
        mean, var = running_mean, running_var
    inv = torch.rsqrt(var + eps)
    return (x - mean.reshape(shape)) * (inv * weight).reshape(shape) + bias.reshape(shape)


# ----------------------------------------------------------------------------------------
# reshuffles / resizes
# ----------------------------------------------------------------------------------------

def pixel_shuffle(x: Tensor, r: int = 2) -> Tensor:
    """nn.PixelShuffle(r): out[b,c,r*h+i,r*w+j] = in[b, c*r*r + r*i + j, h, w] (model/tsrn.py:469)."""
    b, c, h, w = x.shape
    co = c // (r * r)
    x = x.reshape(b, co, r, r, h, w).permute(0, 1, 4, 2, 5, 3)
    return x.reshape(b, co, h * r, w * r)


def _cubic_coeffs(t: float, a: float = -0.75) -> Tuple[float, float, float, float]:
    def c1(x):  # |x| <= 1
        return ((a + 2) * x - (a + 3)) * x * x + 1
    def c2(x):  # 1 < |x| < 2
        return ((a * x - 5 * a) * x + 8 * a) * x - 4 * a
    return c2(t + 1.0), c1(t), c1(1.0 - t), c2(2.0 - t)


def bicubic_matrix(n_in: int, n_out: int) -> Tensor:
    """[n_out, n_in] interpolation matrix of F.interpolate(mode='bicubic', align_corners=False):
    src = (dst+0.5)*n_in/n_out - 0.5, A=-0.75, neighbour indices clamped (interfaces/base.py:822)."""
    m = torch.zeros(n_out, n_in, dtype=torch.float64)
    scale = n_in / n_out
    for o in range(n_out):
        src = (o + 0.5) * scale - 0.5
        i0 = math.floor(src)
        t = src - i0
        for k, wgt in enumerate(_cubic_coeffs(t)):
            idx = min(max(i0 - 1 + k, 0), n_in - 1)
            m[o, idx] += wgt
    return m


def bilinear_ac_matrix(n_in: int, n_out: int) -> Tensor:
    """[n_out, n_in] matrix of F.interpolate(mode='bilinear', align_corners=True) along one axis
    (model/tsrn.py:196)."""
    m = torch.zeros(n_out, n_in, dtype=torch.float64)
    for o in range(n_out):
        src = o * (n_in - 1) / (n_out - 1) if n_out > 1 else 0.0
        i0 = min(int(math.floor(src)), n_in - 1)
        i1 = min(i0 + 1, n_in - 1)
        t = src - i0
        m[o, i0] += 1.0 - t
        m[o, i1] += t
    return m


def resize_separable(x: Tensor, mh: Tensor, mw: Tensor) -> Tensor:
    """Apply per-axis interpolation matrices: out = mh @ x @ mw^T over the last two dims."""
    return torch.einsum('oh,bchw,pw->bcop', mh.to(x.dtype), x, mw.to(x.dtype))


def parse_crnn_data(imgs: Tensor, out_hw=(32, 100)) -> Tensor:
    """interfaces/base.py:806-829 — bicubic resize of RGB to (32,100) then
    0.299 R + 0.587 G + 0.114 B → [B,1,32,100]."""
    x = imgs[:, :3]
    mh = bicubic_matrix(x.shape[2], out_hw[0])
    mw = bicubic_matrix(x.shape[3], out_hw[1])
    y = resize_separable(x, mh, mw)
    return 0.299 * y[:, 0:1] + 0.587 * y[:, 1:2] + 0.114 * y[:, 2:3]


# ----------------------------------------------------------------------------------------
# recurrences
# ----------------------------------------------------------------------------------------

def gru_direction(x: Tensor, w_ih: Tensor, w_hh: Tensor, b_ih: Tensor, b_hh: Tensor, reverse: bool) -> Tensor:
    """One direction of nn.GRU, batch-first x [S,T,I] → [S,T,H]; gate order r,z,n;
    n = tanh(W_in x + b_in + r*(W_hn h + b_hn)); h' = (1-z)*n + z*h."""
    S, T, _ = x.shape
    H = w_hh.shape[1]
    gx = x @ w_ih.t() + b_ih                      # [S,T,3H]
    h = x.new_zeros(S, H)
    outs = [None] * T
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        gh = h @ w_hh.t() + b_hh
        r = torch.sigmoid(gx[:, t, :H] + gh[:, :H])
        z = torch.sigmoid(gx[:, t, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gx[:, t, 2 * H:] + r * gh[:, 2 * H:])
        h = (1 - z) * n + z * h
        outs[t] = h
    return torch.stack(outs, 1)


def bigru(x: Tensor, p: Dict[str, Tensor], prefix: str) -> Tensor:
    """nn.GRU(bidirectional=True, batch_first=True) (model/tsrn.py:496,504)."""
    f = gru_direction(x, p[prefix + 'weight_ih_l0'], p[prefix + 'weight_hh_l0'],
                      p[prefix + 'bias_ih_l0'], p[prefix + 'bias_hh_l0'], False)
    b = gru_direction(x, p[prefix + 'weight_ih_l0_reverse'], p[prefix + 'weight_hh_l0_reverse'],
                      p[prefix + 'bias_ih_l0_reverse'], p[prefix + 'bias_hh_l0_reverse'], True)
    return torch.cat([f, b], -1)


def lstm_direction(x: Tensor, w_ih: Tensor, w_hh: Tensor, b_ih: Tensor, b_hh: Tensor, reverse: bool) -> Tensor:
    """One direction of nn.LSTM, time-first x [T,B,I] → [T,B,H]; gate order i,f,g,o."""
    T, B, _ = x.shape
    H = w_hh.shape[1]
    gx = x @ w_ih.t() + b_ih + b_hh
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    outs = [None] * T
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        g = gx[t] + h @ w_hh.t()
        i = torch.sigmoid(g[:, :H])
        f = torch.sigmoid(g[:, H:2 * H])
        gg = torch.tanh(g[:, 2 * H:3 * H])
        o = torch.sigmoid(g[:, 3 * H:])
        c = f * c + i * gg
        h = o * torch.tanh(c)
        outs[t] = h
    return torch.stack(outs, 0)


def bilstm(x: Tensor, p: Dict[str, Tensor], prefix: str) -> Tensor:
    """nn.LSTM(bidirectional=True), time-first (model/crnn/crnn.py:10,19)."""
    f = lstm_direction(x, p[prefix + 'weight_ih_l0'], p[prefix + 'weight_hh_l0'],
                       p[prefix + 'bias_ih_l0'], p[prefix + 'bias_hh_l0'], False)
    b = lstm_direction(x, p[prefix + 'weight_ih_l0_reverse'], p[prefix + 'weight_hh_l0_reverse'],
                       p[prefix + 'bias_ih_l0_reverse'], p[prefix + 'bias_hh_l0_reverse'], True)
    return torch.cat([f, b], -1)


# ----------------------------------------------------------------------------------------
# TPS rectification + bilinear sampler
# ----------------------------------------------------------------------------------------

def tps_partial_repr(a: Tensor, b: Tensor) -> Tensor:
    """phi(r) = 0.5 r^2 log r^2 with 0 log 0 := 0 (model/tps_spatial_transformer.py:22-34)."""
    d = a[:, None, :] - b[None, :, :]
    d2 = d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]
    out = 0.5 * d2 * torch.log(d2)
    return torch.where(torch.isnan(out), torch.zeros_like(out), out)


def tps_control_points(n: int = 20, margins=(0.05, 0.05)) -> Tensor:
    """model/tps_spatial_transformer.py:38-50 (float64 linspace → float32, like np→torch.Tensor)."""
    import numpy as np
    mx, my = margins
    k = n // 2
    xs = np.linspace(mx, 1.0 - mx, k)
    top = np.stack([xs, np.ones(k) * my], 1)
    bot = np.stack([xs, np.ones(k) * (1.0 - my)], 1)
    return torch.tensor(np.concatenate([top, bot], 0), dtype=torch.float32)


def tps_buffers(h: int, w: int, n: int = 20, margins=(0.05, 0.05)) -> Dict[str, Tensor]:
    """The four registered buffers of TPSSpatialTransformer (model/tps_spatial_transformer.py:56-95)."""
    tcp = tps_control_points(n, margins)
    fk = torch.zeros(n + 3, n + 3)
    fk[:n, :n] = tps_partial_repr(tcp, tcp)
    fk[:n, -3] = 1
    fk[-3, :n] = 1
    fk[:n, -2:] = tcp
    fk[-2:, :n] = tcp.t()
    inv = torch.inverse(fk)
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing='ij')
    coord = torch.stack([xs.reshape(-1) / (w - 1), ys.reshape(-1) / (h - 1)], 1)
    rep = torch.cat([tps_partial_repr(coord, tcp), torch.ones(h * w, 1), coord], 1)
    return {'inverse_kernel': inv, 'padding_matrix': torch.zeros(3, 2),
            'target_coordinate_repr': rep, 'target_control_points': tcp}


def grid_sample_bilinear_zeros(img: Tensor, grid: Tensor) -> Tensor:
    """F.grid_sample(img, grid) with torch-2.11 defaults: bilinear, zeros padding,
    align_corners=False: ix = ((x+1)*W-1)/2 (model/tps_spatial_transformer.py:10-11)."""
    B, C, H, W = img.shape
    _, Ho, Wo, _ = grid.shape
    ix = ((grid[..., 0] + 1) * W - 1) / 2
    iy = ((grid[..., 1] + 1) * H - 1) / 2
    x0 = torch.floor(ix)
    y0 = torch.floor(iy)
    tx = ix - x0
    ty = iy - y0
    flat = img.reshape(B, C, H * W)
    out = img.new_zeros(B, C, Ho, Wo)
    for dy, wy in ((0, 1 - ty), (1, ty)):
        for dx, wx in ((0, 1 - tx), (1, tx)):
            xx = x0 + dx
            yy = y0 + dy
            valid = ((xx >= 0) & (xx <= W - 1) & (yy >= 0) & (yy <= H - 1)).to(img.dtype)
            idx = (yy.clamp(0, H - 1) * W + xx.clamp(0, W - 1)).long().reshape(B, 1, Ho * Wo).expand(B, C, Ho * Wo)
            v = torch.gather(flat, 2, idx).reshape(B, C, Ho, Wo)
            out = out + v * (wy * wx * valid).unsqueeze(1)
    return out


def tps_forward(img: Tensor, ctrl: Tensor, buf: Dict[str, Tensor], h: int, w: int) -> Tuple[Tensor, Tensor]:
    """TPSSpatialTransformer.forward (model/tps_spatial_transformer.py:97-112)."""
    B = ctrl.shape[0]
    Y = torch.cat([ctrl, buf['padding_matrix'].expand(B, 3, 2)], 1)
    M = torch.matmul(buf['inverse_kernel'], Y)
    src = torch.matmul(buf['target_coordinate_repr'], M)
    grid = src.reshape(-1, h, w, 2).clamp(0, 1) * 2.0 - 1.0
    return grid_sample_bilinear_zeros(img, grid), src


# ----------------------------------------------------------------------------------------
# losses and metrics
# ----------------------------------------------------------------------------------------

def gradient_map(x: Tensor) -> Tensor:
    """GradientPriorLoss.gradient_map (loss/image_loss.py:43-51): central differences with zero padding."""
    r = F.pad(x, (0, 1, 0, 0))[:, :, :, 1:]
    l = F.pad(x, (1, 0, 0, 0))[:, :, :, :-1]
    t = F.pad(x, (0, 0, 1, 0))[:, :, :-1, :]
    b = F.pad(x, (0, 0, 0, 1))[:, :, 1:, :]
    return torch.sqrt(((r - l) * 0.5) ** 2 + ((t - b) * 0.5) ** 2 + 1e-6)


def image_loss(out: Tensor, tgt: Tensor, gradient: bool = True, loss_weight=(1.0, 1e-4)) -> Tensor:
    """ImageLoss.forward (loss/image_loss.py:19-30)."""
    mse = ((out - tgt) ** 2).mean()
    if not gradient:
        return loss_weight[0] * mse
    gp = (gradient_map(out[:, :3]) - gradient_map(tgt[:, :3])).abs().mean()
    return loss_weight[0] * mse + loss_weight[1] * gp


def semantic_loss(pred: Tensor, gt: Tensor) -> Tensor:
    """SemanticLoss.forward (loss/semantic_loss.py:21-38); KLDivLoss default reduction 'mean' = element mean
    of target*(log target - input)."""
    margin = (gt - pred).abs().mean()
    inp = torch.log(pred + 1e-20)
    tg = gt + 1e-20
    kl = (tg * (torch.log(tg) - inp)).mean()
    return margin + kl


def calculate_psnr(a: Tensor, b: Tensor) -> Tensor:
    """utils/ssim_psnr.py:9-15."""
    mse = ((a[:, :3] * 255 - b[:, :3] * 255) ** 2).mean()
    return 20 * torch.log10(255.0 / torch.sqrt(mse))


def ssim(a: Tensor, b: Tensor, window_size: int = 11, sigma: float = 1.5) -> Tensor:
    """utils/ssim_psnr.py:18-78: 11x11 gaussian, zero-padded depthwise conv, on RGB, global mean."""
    a = a[:, :3]
    b = b[:, :3]
    c = a.shape[1]
    g = torch.tensor([math.exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)])
    g = g / g.sum()
    win = (g[:, None] @ g[None, :]).float().expand(c, 1, window_size, window_size).contiguous()
    pad = window_size // 2
    conv = lambda t: F.conv2d(t, win, padding=pad, groups=c)
    mu1, mu2 = conv(a), conv(b)
    s1 = conv(a * a) - mu1 * mu1
    s2 = conv(b * b) - mu2 * mu2
    s12 = conv(a * b) - mu1 * mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s1 + s2 + C2))
    return m.mean()


# ----------------------------------------------------------------------------------------
# optimiser step glue (interfaces/super_resolution.py:419-424, interfaces/base.py:449-450)
# ----------------------------------------------------------------------------------------

def clip_grad_norm(grads, max_norm: float = 0.25) -> Tensor:
    """torch.nn.utils.clip_grad_norm_: global L2, scale by max_norm/(norm+1e-6) clamped to 1, in place."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in grads:
        g.mul_(coef)
    return total


def adam_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float = 1e-3,
              beta1: float = 0.5, beta2: float = 0.999, eps: float = 1e-8) -> None:
    """torch.optim.Adam (no weight decay, no amsgrad), in place; ``step`` is 1-based."""
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)
