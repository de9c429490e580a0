"""TPS spatial transformer — buffer container mirroring ``model/tps_spatial_transformer.py:52-95``.

Holds the reference's four registered buffers (same names / shapes, so reference checkpoints load) plus a transposed copy
``reprT`` [23, H*W] that the fused TPS+bilinear-sampler kernel reads coalesced (``csrc/tps.cu``).
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn


def _partial_repr(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    # phi(r) = 0.5 * r^2 * log(r^2), 0*log(0) := 0
    d = a.view(-1, 1, 2) - b.view(1, -1, 2)
    d2 = d[:, :, 0] * d[:, :, 0] + d[:, :, 1] * d[:, :, 1]
    out = 0.5 * d2 * torch.log(d2)
    out.masked_fill_(out != out, 0)
    return out


def _target_control_points(n: int, margins) -> torch.Tensor:
    mx, my = margins
    k = n // 2
    xs = np.linspace(mx, 1.0 - mx, k)
    top = np.stack([xs, np.ones(k) * my], axis=1)
    bot = np.stack([xs, np.ones(k) * (1.0 - my)], axis=1)
    return torch.Tensor(np.concatenate([top, bot], axis=0))


class TPSSpatialTransformer(nn.Module):
    def __init__(self, output_image_size=None, num_control_points=None, margins=None):
        super().__init__()
        self.output_image_size, self.num_control_points, self.margins = output_image_size, num_control_points, margins
        self.target_height, self.target_width = output_image_size
        H, W, N = self.target_height, self.target_width, num_control_points
        tcp = _target_control_points(N, margins)
        fk = torch.zeros(N + 3, N + 3)
        fk[:N, :N].copy_(_partial_repr(tcp, tcp))
        fk[:N, -3].fill_(1)
        fk[-3, :N].fill_(1)
        fk[:N, -2:].copy_(tcp)
        fk[-2:, :N].copy_(tcp.transpose(0, 1))
        inverse_kernel = torch.inverse(fk)
        ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing='ij')
        coord = torch.stack([xs.reshape(-1) / (W - 1), ys.reshape(-1) / (H - 1)], dim=1)
        rep = torch.cat([_partial_repr(coord, tcp), torch.ones(H * W, 1), coord], dim=1)
        self.register_buffer('inverse_kernel', inverse_kernel)
        self.register_buffer('padding_matrix', torch.zeros(3, 2))
        self.register_buffer('target_coordinate_repr', rep)
        self.register_buffer('target_control_points', tcp)
        self.register_buffer('reprT', rep.t().contiguous(), persistent=False)

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        # in place: captured CUDA graphs (train steps) hold this buffer's address
        rt = self.target_coordinate_repr.t()
        if self.reprT.shape == rt.shape and self.reprT.device == rt.device:
            self.reprT.copy_(rt)
        else:
            self.reprT = rt.contiguous()

    def forward(self, input, source_control_points):
        raise RuntimeError('TPSSpatialTransformer is executed inside TSRN/TSRN_TL.forward by the tpgsr_b200 CUDA path')
