"""Oracle model forwards: functional restatements driven by a reference-keyed state dict.

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.  ``p`` is always a dict with the
reference's ``state_dict`` keys; BN running stats in ``p`` are updated in
place in training mode, as the reference's modules do.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import ops

Tensor = torch.Tensor
P = Dict[str, Tensor]


def _bn(x: Tensor, p: P, pre: str, training: bool) -> Tensor:
    y = ops.batchnorm(x, p[pre + 'weight'], p[pre + 'bias'], p[pre + 'running_mean'], p[pre + 'running_var'], training)
    if training and (pre + 'num_batches_tracked') in p:
        p[pre + 'num_batches_tracked'] += 1
    return y


def stn_head(x: Tensor, p: P, pre: str = 'stn_head.', training: bool = True):
    """STNHead.forward (model/stn_head.py:92-106; net :34-57)."""
    pools = {0: (2, 2), 2: (2, 2), 4: (2, 2), 6: (2, 2), 8: (1, 2)}
    for i in (0, 2, 4, 6, 8, 10):
        q = f'{pre}stn_convnet.{i}.'
        x = F.conv2d(x, p[q + '0.weight'], p[q + '0.bias'], padding=1)
        x = torch.relu(_bn(x, p, q + '1.', training))
        if i in pools:
            x = F.max_pool2d(x, pools[i], pools[i])
    x = x.reshape(x.shape[0], -1)
    f = x @ p[pre + 'stn_fc1.0.weight'].t() + p[pre + 'stn_fc1.0.bias']
    f = torch.relu(_bn(f, p, pre + 'stn_fc1.1.', training))
    c = (0.1 * f) @ p[pre + 'stn_fc2.weight'].t() + p[pre + 'stn_fc2.bias']
    return f, c.reshape(-1, p[pre + 'stn_fc2.bias'].numel() // 2, 2)


def gru_block(x: Tensor, p: P, pre: str) -> Tensor:
    """GruBlock.forward (model/tsrn.py:491-508): 1x1 conv, rows of the map are sequences over W."""
    x = F.conv2d(x, p[pre + 'conv1.weight'], p[pre + 'conv1.bias'])
    b, c, h, w = x.shape
    seq = x.permute(0, 2, 3, 1).reshape(b * h, w, c)
    y = ops.bigru(seq, p, pre + 'gru.')
    return y.reshape(b, h, w, c).permute(0, 3, 1, 2)


def srb(x: Tensor, p: P, pre: str, training: bool, text: Optional[Tensor] = None) -> Tensor:
    """RecurrentResidualBlock / RecurrentResidualBlockTL (model/tsrn.py:373-426)."""
    r = F.conv2d(x, p[pre + 'conv1.weight'], p[pre + 'conv1.bias'], padding=1)
    r = ops.mish(_bn(r, p, pre + 'bn1.', training))
    r = F.conv2d(r, p[pre + 'conv2.weight'], p[pre + 'conv2.bias'], padding=1)
    r = _bn(r, p, pre + 'bn2.', training)
    if text is not None:
        r = torch.cat([r, text], 1)
    r = gru_block(r.transpose(-1, -2), p, pre + 'gru1.').transpose(-1, -2)
    return gru_block(x + r, p, pre + 'gru2.')


def info_gen(t: Tensor, p: P, pre: str, training: bool, variant: str = 'tsrn') -> Tensor:
    """InfoGen.forward (model/tsrn.py:81-108; srresnet variant model/srresnet.py:166-195)."""
    if variant == 'tsrn':
        cfg = [((2, 2), (1, 1)), ((2, 2), (1, 1)), ((2, 2), (1, 1)), ((2, 1), (1, 0))]
    else:  # srresnet.py: tconv1/2 have no padding, tconv4 padding=1
        cfg = [((2, 2), (0, 0)), ((2, 2), (0, 0)), ((2, 2), (1, 1)), ((2, 1), (1, 1))]
    x = t
    for i, (s, pd) in enumerate(cfg, 1):
        x = F.conv_transpose2d(x, p[f'{pre}tconv{i}.weight'], None, stride=s, padding=pd)
        x = torch.relu(_bn(x, p, f'{pre}bn{i}.', training))
    return x


def tsrn(x: Tensor, p: P, training: bool, stn: bool = True, srb_nums: int = 5,
         text_emb: Optional[Tensor] = None, tl: bool = False, return_aux: bool = False):
    """TSRN.forward (model/tsrn.py:61-78) and TSRN_TL.forward (model/tsrn.py:178-215)."""
    aux = {}
    if stn and training:
        _, ctrl = stn_head(x, p, 'stn_head.', training)
        buf = {k[4:]: v for k, v in p.items() if k.startswith('tps.')}
        x, src = ops.tps_forward(x, ctrl, buf, x.shape[2], x.shape[3])
        aux['ctrl'] = ctrl
        aux['rectified'] = x
    b1 = ops.prelu(F.conv2d(x, p['block1.0.weight'], p['block1.0.bias'], padding=4), p['block1.1.weight'])
    text = None
    if tl:
        if text_emb is None:
            text_emb = x.new_zeros(x.shape[0], p['infoGen.tconv1.weight'].shape[0], 1, 26)
        t = info_gen(text_emb, p, 'infoGen.', training)
        mh = ops.bilinear_ac_matrix(t.shape[2], x.shape[2])
        mw = ops.bilinear_ac_matrix(t.shape[3], x.shape[3])
        text = ops.resize_separable(t, mh, mw)
        aux['text'] = text
    b = b1
    for i in range(srb_nums):
        b = srb(b, p, f'block{i + 2}.', training, text)
    k = srb_nums + 2
    b = F.conv2d(b, p[f'block{k}.0.weight'], p[f'block{k}.0.bias'], padding=1)
    b = _bn(b, p, f'block{k}.1.', training)
    k += 1
    u = F.conv2d(b1 + b, p[f'block{k}.0.conv.weight'], p[f'block{k}.0.conv.bias'], padding=1)
    u = ops.mish(ops.pixel_shuffle(u, 2))
    out = torch.tanh(F.conv2d(u, p[f'block{k}.1.weight'], p[f'block{k}.1.bias'], padding=4))
    return (out, aux) if return_aux else out


def srresnet_tl(x: Tensor, p: P, training: bool, text_emb: Optional[Tensor] = None, return_aux: bool = False):
    """SRResNet_TL.forward, STN off (model/srresnet.py:139-163; ResidualBlock_TL :198-220; UpsampleBLock :223-234; InfoGen variant
    :166-195 — tconv1/2 without padding, so the prior map is [B,32,25,213] before the bilinear(align_corners=True) resize)."""
    if text_emb is None:
        text_emb = x.new_zeros(x.shape[0], p['infoGen.tconv1.weight'].shape[0], 1, 26)
    t = info_gen(text_emb, p, 'infoGen.', training, variant='srresnet')
    text = ops.resize_separable(t, ops.bilinear_ac_matrix(t.shape[2], x.shape[2]), ops.bilinear_ac_matrix(t.shape[3], x.shape[3]))
    b1 = ops.prelu(F.conv2d(x, p['block1.0.weight'], p['block1.0.bias'], padding=4), p['block1.1.weight'])
    b = b1
    for i in range(2, 7):
        q = f'block{i}.'
        r = F.conv2d(b, p[q + 'conv1.weight'], p[q + 'conv1.bias'], padding=1)
        r = ops.prelu(_bn(r, p, q + 'bn1.', training), p[q + 'prelu.weight'])
        r = F.conv2d(torch.cat([r, text], 1), p[q + 'conv2.weight'], p[q + 'conv2.bias'], padding=1)
        b = b + _bn(r, p, q + 'bn2.', training)
    b7 = _bn(F.conv2d(b, p['block7.0.weight'], p['block7.0.bias'], padding=1), p, 'block7.1.', training)
    u = F.conv2d(b1 + b7, p['block8.0.conv.weight'], p['block8.0.conv.bias'], padding=1)
    u = ops.prelu(ops.pixel_shuffle(u, 2), p['block8.0.prelu.weight'])
    out = torch.tanh(F.conv2d(u, p['block8.1.weight'], p['block8.1.bias'], padding=4))
    return (out, {'text': text, 't': t}) if return_aux else out


def crnn(x: Tensor, p: P, training: bool) -> Tensor:
    """CRNN.forward (model/crnn/crnn.py:29-90) → logits [26,B,37]."""
    def conv(i, x, pad):
        return F.conv2d(x, p[f'cnn.conv{i}.weight'], p[f'cnn.conv{i}.bias'], padding=pad)
    x = F.max_pool2d(torch.relu(conv(0, x, 1)), 2, 2)
    x = F.max_pool2d(torch.relu(conv(1, x, 1)), 2, 2)
    x = torch.relu(_bn(conv(2, x, 1), p, 'cnn.batchnorm2.', training))
    x = F.max_pool2d(torch.relu(conv(3, x, 1)), (2, 2), (2, 1), (0, 1))
    x = torch.relu(_bn(conv(4, x, 1), p, 'cnn.batchnorm4.', training))
    x = F.max_pool2d(torch.relu(conv(5, x, 1)), (2, 2), (2, 1), (0, 1))
    x = torch.relu(_bn(conv(6, x, 0), p, 'cnn.batchnorm6.', training))
    assert x.shape[2] == 1, 'the height of conv must be 1'
    seq = x.squeeze(2).permute(2, 0, 1)
    for i in (0, 1):
        r = ops.bilstm(seq, p, f'rnn.{i}.rnn.')
        T, B, Hh = r.shape
        seq = (r.reshape(T * B, Hh) @ p[f'rnn.{i}.embedding.weight'].t() + p[f'rnn.{i}.embedding.bias']).reshape(T, B, -1)
    return seq


def srcnn(x: Tensor, p: P) -> Tensor:
    """SRCNN.forward, STN=False (model/srcnn.py:109-145): nearest x2 → 9x9+ReLU → 1x1+ReLU → 5x5."""
    x = x.repeat_interleave(2, 2).repeat_interleave(2, 3)
    x = torch.relu(F.conv2d(x, p['conv1.weight'], p['conv1.bias'], padding=4))
    x = torch.relu(F.conv2d(x, p['conv2.weight'], p['conv2.bias']))
    return F.conv2d(x, p['conv3.weight'], p['conv3.bias'], padding=2)


def prior_from_logits(logits: Tensor, drop_first_quarter: bool = True) -> Tensor:
    """interfaces/super_resolution.py:319-321,376-382: softmax → [B,37,1,26] → zero the first B//4 samples.
    Returns (probs[T,B,C], prior[B,C,1,T])."""
    pr = torch.softmax(logits, -1)
    tp = pr.permute(1, 0, 2).unsqueeze(1).permute(0, 3, 1, 2)
    if drop_first_quarter:
        drop = torch.ones(tp.shape[0])
        drop[: tp.shape[0] // 4] = 0.
        tp = tp * drop.reshape(-1, 1, 1, 1)
    return pr, tp
