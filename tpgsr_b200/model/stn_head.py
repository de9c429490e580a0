"""STN localisation head — parameter container mirroring ``model/stn_head.py:25-90`` (state_dict keys, initialisers).

The arithmetic (6x conv3x3+BN+ReLU with max-pools, fc+BN1d+ReLU, fc on 0.1*feat) runs in
``tpgsr_b200.engine.stn_head_fwd/bwd`` on the CUDA library; this class only owns the parameters.
"""
from __future__ import annotations

import math

import numpy as np
import torch
from torch import nn


def _conv_block(cin: int, cout: int) -> nn.Sequential:
    return nn.Sequential(nn.Conv2d(cin, cout, kernel_size=3, stride=1, padding=1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class STNHead(nn.Module):
    def __init__(self, in_planes, num_ctrlpoints, activation='none', input_size=(32, 128)):
        super().__init__()
        if activation != 'none':
            raise NotImplementedError('tpgsr_b200 STNHead supports activation="none" (the only mode TSRN uses)')
        self.in_planes, self.num_ctrlpoints, self.activation = in_planes, num_ctrlpoints, activation
        chans = [in_planes, 32, 64, 128, 256, 256, 256]
        pools = [(2, 2)] * 4 + [(1, 2)]
        layers = []
        for i in range(6):
            layers.append(_conv_block(chans[i], chans[i + 1]))
            if i < 5:
                layers.append(nn.MaxPool2d(kernel_size=pools[i], stride=pools[i]))
        self.stn_convnet = nn.Sequential(*layers)
        self.stn_fc1 = nn.Sequential(nn.Linear(2 * 256, 512), nn.BatchNorm1d(512), nn.ReLU(inplace=True))
        self.stn_fc2 = nn.Linear(512, num_ctrlpoints * 2)
        self._init_like_reference()

    def _init_like_reference(self):
        # model/stn_head.py:59-90: He-style normal for convs (fan = k*k*Cout), N(0, 1e-3) for fc1, fc2 = constant control points
        for m in list(self.stn_convnet.modules()) + list(self.stn_fc1.modules()):
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2. / n))
                m.bias.data.zero_()
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
            elif isinstance(m, nn.Linear):
                m.weight.data.normal_(0, 0.001)
                m.bias.data.zero_()
        margin = 0.01
        k = self.num_ctrlpoints // 2
        xs = np.linspace(margin, 1. - margin, k)
        top = np.stack([xs, np.ones(k) * margin], axis=1)
        bot = np.stack([xs, np.ones(k) * (1 - margin)], axis=1)
        pts = np.concatenate([top, bot], axis=0).astype(np.float32)
        self.stn_fc2.weight.data.zero_()
        self.stn_fc2.bias.data = torch.Tensor(pts).view(-1)

    def forward(self, x):
        raise RuntimeError('STNHead is executed inside TSRN/TSRN_TL.forward by the tpgsr_b200 CUDA path')
