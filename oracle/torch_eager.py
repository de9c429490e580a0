"""Stock-PyTorch (ATen / cuDNN / cuBLAS) statement of the hot path — MEASUREMENT INFRASTRUCTURE ONLY.

``bench.py``'s developer leg ``gpu_eager_baseline`` times this on the same B200, same batches, TF32 off: it is "what the reference's
own modules cost when simply moved to this GPU"  — ``nn.Conv2d`` / ``nn.BatchNorm2d`` / ``nn.GRU`` /
``nn.LSTM`` (cuDNN RNN) / ``F.grid_sample`` / ``F.interpolate`` / ``nn.PixelShuffle`` / autograd / ``clip_grad_norm_`` /
``torch.optim.Adam``.  Never imported by ``tpgsr_b200/``; never the thing whose number is reported as ``value``.

The modules carry the reference's ``state_dict`` keys, so weights load straight from the drop-in modules.
Semantics restated from the same call sites as ``oracle/models.py`` (model/tsrn.py:18-215,373-508; model/stn_head.py:34-106;
model/tps_spatial_transformer.py:97-112; model/crnn/crnn.py:5-90; loss/image_loss.py:19-51; loss/semantic_loss.py:21-38;
interfaces/super_resolution.py:297-424); ``tests/test_oracle_golden.py::test_torch_eager_matches_oracle`` pins it to the oracle on CPU.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn.functional as F
from torch import nn

from . import ops as o


def _mish(x):
    return x * torch.tanh(F.softplus(x))


def _cbr(ci, co):
    return nn.Sequential(nn.Conv2d(ci, co, 3, 1, 1), nn.BatchNorm2d(co), nn.ReLU(inplace=True))


class _GruBlock(nn.Module):
    def __init__(self, ci, co):
        super().__init__()
        self.conv1 = nn.Conv2d(ci, co, 1)
        self.gru = nn.GRU(co, co // 2, bidirectional=True, batch_first=True)

    def forward(self, x):
        x = self.conv1(x).permute(0, 2, 3, 1).contiguous()
        b, h, w, c = x.shape
        y, _ = self.gru(x.view(b * h, w, c))
        return y.view(b, h, w, c).permute(0, 3, 1, 2)


class _SRB(nn.Module):
    def __init__(self, ch, text_ch=0):
        super().__init__()
        self.conv1, self.bn1 = nn.Conv2d(ch, ch, 3, padding=1), nn.BatchNorm2d(ch)
        self.conv2, self.bn2 = nn.Conv2d(ch, ch, 3, padding=1), nn.BatchNorm2d(ch)
        self.gru1, self.gru2 = _GruBlock(ch + text_ch, ch), _GruBlock(ch, ch)

    def forward(self, x, text=None):
        r = self.bn2(self.conv2(_mish(self.bn1(self.conv1(x)))))
        if text is not None:
            r = torch.cat([r, text], 1)
        r = self.gru1(r.transpose(-1, -2)).transpose(-1, -2)
        return self.gru2(x + r).contiguous()


class _Up(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch * 4, 3, padding=1)

    def forward(self, x):
        return _mish(F.pixel_shuffle(self.conv(x), 2))


class _STNHead(nn.Module):
    def __init__(self):
        super().__init__()
        chans, layers = [4, 32, 64, 128, 256, 256, 256], []
        for i in range(6):
            layers.append(_cbr(chans[i], chans[i + 1]))
            if i < 5:
                layers.append(nn.MaxPool2d(2, 2) if i < 4 else nn.MaxPool2d((1, 2), (1, 2)))
        self.stn_convnet = nn.Sequential(*layers)
        self.stn_fc1 = nn.Sequential(nn.Linear(512, 512), nn.BatchNorm1d(512), nn.ReLU(inplace=True))
        self.stn_fc2 = nn.Linear(512, 40)

    def forward(self, x):
        f = self.stn_fc1(self.stn_convnet(x).flatten(1))
        return self.stn_fc2(0.1 * f).view(-1, 20, 2)


class _TPS(nn.Module):
    def __init__(self, h=16, w=64):
        super().__init__()
        self.hw = (h, w)
        for k, v in o.tps_buffers(h, w).items():
            self.register_buffer(k, v)

    def forward(self, img, ctrl):
        y = torch.cat([ctrl, self.padding_matrix.expand(ctrl.shape[0], 3, 2)], 1)
        src = self.target_coordinate_repr @ (self.inverse_kernel @ y)
        grid = src.view(-1, *self.hw, 2).clamp(0, 1) * 2 - 1
        return F.grid_sample(img, grid, mode='bilinear', padding_mode='zeros', align_corners=False)


class _InfoGen(nn.Module):
    def __init__(self, t_emb=37, out=32):
        super().__init__()
        cfg = [(t_emb, 512, 2, 1), (512, 128, 2, 1), (128, 64, 2, 1), (64, out, (2, 1), (1, 0))]
        for i, (ci, co, s, p) in enumerate(cfg, 1):
            setattr(self, f'tconv{i}', nn.ConvTranspose2d(ci, co, 3, s, p, bias=False))
            setattr(self, f'bn{i}', nn.BatchNorm2d(co))

    def forward(self, t):
        for i in range(1, 5):
            t = F.relu(getattr(self, f'bn{i}')(getattr(self, f'tconv{i}')(t)))
        return t


class EagerTSRN(nn.Module):
    """TSRN (tl=False) / TSRN_TL (tl=True), STN + mask, 5 SRBs, 64 channels."""

    def __init__(self, tl: bool = False):
        super().__init__()
        self.tl = tl
        self.block1 = nn.Sequential(nn.Conv2d(4, 64, 9, padding=4), nn.PReLU())
        for i in range(2, 7):
            setattr(self, f'block{i}', _SRB(64, 32 if tl else 0))
        self.block7 = nn.Sequential(nn.Conv2d(64, 64, 3, padding=1), nn.BatchNorm2d(64))
        self.block8 = nn.Sequential(_Up(64), nn.Conv2d(64, 4, 9, padding=4))
        self.stn_head, self.tps = _STNHead(), _TPS()
        if tl:
            self.infoGen = _InfoGen()

    def forward(self, x, text_emb: Optional[torch.Tensor] = None):
        if self.training:
            x = self.tps(x, self.stn_head(x))
        b1 = self.block1(x)
        text = None
        if self.tl:
            text = F.interpolate(self.infoGen(text_emb), (x.shape[2], x.shape[3]), mode='bilinear', align_corners=True)
        b = b1
        for i in range(2, 7):
            b = getattr(self, f'block{i}')(b, text)
        return torch.tanh(self.block8(b1 + self.block7(b)))


class _BiLSTM(nn.Module):
    def __init__(self, n_in, hid, n_out):
        super().__init__()
        self.rnn = nn.LSTM(n_in, hid, bidirectional=True)
        self.embedding = nn.Linear(2 * hid, n_out)

    def forward(self, x):
        r, _ = self.rnn(x)
        t, b, h = r.shape
        return self.embedding(r.view(t * b, h)).view(t, b, -1)


class EagerCRNN(nn.Module):
    def __init__(self):
        super().__init__()
        ks, ps, nm = [3, 3, 3, 3, 3, 3, 2], [1, 1, 1, 1, 1, 1, 0], [64, 128, 256, 256, 512, 512, 512]
        cnn = nn.Sequential()
        for i in range(7):
            cnn.add_module(f'conv{i}', nn.Conv2d(1 if i == 0 else nm[i - 1], nm[i], ks[i], 1, ps[i]))
            if i in (2, 4, 6):
                cnn.add_module(f'batchnorm{i}', nn.BatchNorm2d(nm[i]))
            cnn.add_module(f'relu{i}', nn.ReLU(True))
            if i in (0, 1):
                cnn.add_module(f'pooling{i}', nn.MaxPool2d(2, 2))
            if i in (3, 5):
                cnn.add_module(f'pooling{i}', nn.MaxPool2d((2, 2), (2, 1), (0, 1)))
        self.cnn = cnn
        self.rnn = nn.Sequential(_BiLSTM(512, 256, 256), _BiLSTM(256, 256, 37))

    def forward(self, x):
        return self.rnn(self.cnn(x).squeeze(2).permute(2, 0, 1))


def parse_crnn_data(imgs):
    y = F.interpolate(imgs[:, :3], (32, 100), mode='bicubic')
    return 0.299 * y[:, 0:1] + 0.587 * y[:, 1:2] + 0.114 * y[:, 2:3]


def _gradient_map(x):
    r, l = F.pad(x, (0, 1, 0, 0))[:, :, :, 1:], F.pad(x, (1, 0, 0, 0))[:, :, :, :-1]
    t, b = F.pad(x, (0, 0, 1, 0))[:, :, :-1, :], F.pad(x, (0, 0, 0, 1))[:, :, 1:, :]
    return torch.sqrt(((r - l) * 0.5) ** 2 + ((t - b) * 0.5) ** 2 + 1e-6)


def image_loss(out, tgt):
    return F.mse_loss(out, tgt) + 1e-4 * F.l1_loss(_gradient_map(out[:, :3]), _gradient_map(tgt[:, :3]))


def semantic_loss(pred, gt):
    return (gt - pred).abs().mean() + F.kl_div(torch.log(pred + 1e-20), gt + 1e-20, reduction='mean')


class EagerStep:
    """One training step of config 2 (no CRNNs) or configs 3/4 (cascade) with stock torch; mirrors tpgsr_b200.train."""

    def __init__(self, sr: EagerTSRN, students: List[EagerCRNN] = (), teacher: Optional[EagerCRNN] = None, stu_iter: int = 1):
        self.sr, self.students, self.teacher, self.stu_iter = sr, list(students), teacher, stu_iter
        params = list(sr.parameters()) + [p for s in self.students for p in s.parameters()]
        self.opt = torch.optim.Adam(params, lr=1e-3, betas=(0.5, 0.999))

    def __call__(self, lr_img, hr_img):
        self.opt.zero_grad(set_to_none=True)
        if self.teacher is None:
            loss = image_loss(self.sr(lr_img), hr_img) * 100
        else:
            with torch.no_grad():
                label_hr = torch.softmax(self.teacher(parse_crnn_data(hr_img)), -1)
            cascade, loss = lr_img, 0.
            for i in range(self.stu_iter):
                probs = torch.softmax(self.students[i if len(self.students) > 1 else 0](parse_crnn_data(cascade)), -1)
                prior = probs.permute(1, 2, 0).unsqueeze(2)
                keep = torch.ones(prior.shape[0], 1, 1, 1, device=prior.device)
                keep[: prior.shape[0] // 4] = 0
                loss = loss + semantic_loss(probs, label_hr) * 100
                cascade = self.sr(lr_img, prior * keep)
                loss = loss + image_loss(cascade, hr_img) * 100
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.sr.parameters(), 0.25)
        self.opt.step()
        return loss.detach()


@torch.no_grad()
def infer(sr: EagerTSRN, students: List[EagerCRNN], lr_img, stu_iter: int = 1):
    cascade = lr_img
    for i in range(stu_iter):
        probs = torch.softmax(students[i if len(students) > 1 else 0](parse_crnn_data(cascade)), -1)
        cascade = sr(lr_img, probs.permute(1, 2, 0).unsqueeze(2))
    return cascade
