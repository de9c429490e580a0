"""Shadow of the third-party ``ptflops`` package for the reference's complexity printout (``interfaces/base.py:372-381,638,684,764``).

The reference calls ``ptflops.get_model_complexity_info(model, (4, 16, 64), ...)`` on the freshly constructed CPU module; ptflops hooks
``nn.Conv2d`` & co. and runs a batch-1 CPU forward.  The tpgsr_b200 drop-ins execute their submodules inside fused CUDA kernels — their
``nn.Conv2d`` children never run, and there is no CPU forward (no fallback, by design).  For those modules the count is ANALYTIC
(``module.complexity_info(input_res)``: multiply-accumulates of the eval-mode forward from the layer shapes — the same number ptflops
prints for the reference's module); any other module (the reference's own out-of-scope architectures) is counted the ptflops way, with
forward hooks on a CPU batch-1 forward.
"""
from __future__ import annotations

import sys
from typing import Tuple

import torch
from torch import nn


def flops_to_string(macs: float, precision: int = 2) -> str:
    for unit, div in (('GMac', 1e9), ('MMac', 1e6), ('KMac', 1e3)):
        if macs >= div:
            return f'{round(macs / div, precision)} {unit}'
    return f'{macs} Mac'


def params_to_string(n: float, precision: int = 2) -> str:
    for unit, div in (('M', 1e6), ('k', 1e3)):
        if n >= div:
            return f'{round(n / div, precision)} {unit}'
    return str(n)


def _hook_count(model: nn.Module, input_res: Tuple[int, ...]) -> float:
    macs = [0.0]

    def conv_hook(m, inp, out):
        k = m.kernel_size[0] * m.kernel_size[1] * (m.in_channels // m.groups)
        macs[0] += out.numel() * k if not isinstance(m, nn.ConvTranspose2d) else inp[0].numel() * m.kernel_size[0] * m.kernel_size[1] * m.out_channels // m.groups

    def lin_hook(m, inp, out):
        macs[0] += out.numel() * m.in_features

    handles = []
    for m in model.modules():
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
            handles.append(m.register_forward_hook(conv_hook))
        elif isinstance(m, nn.Linear):
            handles.append(m.register_forward_hook(lin_hook))
    was_training = model.training
    model.eval()
    try:
        with torch.no_grad():
            p = next(model.parameters(), None)
            x = torch.zeros((1,) + tuple(input_res), device=p.device if p is not None else 'cpu')
            model(x)
    finally:
        for h in handles:
            h.remove()
        model.train(was_training)
    return macs[0]


def get_model_complexity_info(model, input_res, print_per_layer_stat=True, as_strings=True, input_constructor=None, ost=sys.stdout,
                              verbose=False, ignore_modules=(), custom_modules_hooks=None, **_unused):
    params = sum(p.numel() for p in model.parameters() if p.requires_grad)
    if hasattr(model, 'complexity_info'):
        macs = float(model.complexity_info(tuple(input_res)))
        if print_per_layer_stat:
            print(f'{type(model).__name__}: {flops_to_string(macs)} (analytic, eval-mode forward, batch 1), {params_to_string(params)} parameters', file=ost)
    else:
        macs = _hook_count(model, tuple(input_res))
    if as_strings:
        return flops_to_string(macs), params_to_string(params)
    return macs, params
