"""Flat parameter / gradient storage behind reference-shaped nn.Modules.

The drop-in modules keep the reference's submodule tree (so ``state_dict`` keys, default initialisers, ``.to()``,
``.train()``, ``.parameters()`` behave exactly like ``model/tsrn.py`` etc.), but every parameter's storage is a view
of ONE flat fp32 buffer, and gradients are views of a second flat buffer.  That gives (a) one NCCL all-reduce per step
on a single buffer, (b) one fused clip+Adam kernel, (c) adjacent forward/reverse RNN weights so both
directions' input projections are a single GEMM.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import nn


class FlatState:
    def __init__(self, module: nn.Module, fused: Sequence[Tuple[str, Sequence[str], Tuple[int, ...]]] = ()):
        """``fused``: (fused_name, [param names to place adjacently, in order], shape of the fused view)."""
        self.module = module
        self.fused = list(fused)
        self.flat = None
        self.gflat = None
        self.views: Dict[str, torch.Tensor] = {}
        self.gviews: Dict[str, torch.Tensor] = {}
        self.offsets: Dict[str, Tuple[int, int]] = {}
        self.nbt = None
        self.arena = None             # the FlatArena this state lives in, if any
        self.frozen = False           # set by the train steps once CUDA graphs hold this buffer's addresses

    # -- layout -----------------------------------------------------------------------------------------------------
    def _order(self, params: Dict[str, nn.Parameter]) -> List[str]:
        placed, order = set(), []
        for _, names, _ in self.fused:
            for n in names:
                order.append(n)
                placed.add(n)
        order += [n for n in params if n not in placed]
        return order

    def is_current(self) -> bool:
        if self.flat is None:
            return False
        # EVERY parameter is checked (a few hundred pointer compares): ``module._apply`` (.to / .cuda / .float) re-runs
        # ``RNNBase.flatten_parameters`` on torch 2.11, which moves the GRU / LSTM weights into a fresh cuDNN buffer while the conv weights
        # keep their flat-buffer storage — looking at the first parameter only would miss that and keep training the stale flat copy
        base, dev = self.flat.data_ptr(), self.flat.device
        n = 0
        for name, p in self.module.named_parameters():
            ent = self.offsets.get(name)
            if ent is None or p.device != dev or p.data_ptr() != base + 4 * ent[0]:
                return False
            n += 1
        return n == len(self.offsets)

    def assert_attached(self) -> None:
        """Per-step guard of the train steps (two pointer compares): the first parameter in flat order is an RNN weight whenever the module
        has one (fused groups are laid out first) — the tensors ``flatten_parameters`` moves — and the last one is a conv / BN parameter."""
        for name, p in self._sentinels:
            if p.data_ptr() != self.flat.data_ptr() + 4 * self.offsets[name][0]:
                raise RuntimeError('parameter %s of %s no longer lives in the flat buffer the captured CUDA graphs use (a .to()/.cuda()/.float() '
                                   'after the train step was built?) — rebuild the train step' % (name, type(self.module).__name__))

    def padded_numel(self) -> int:
        return (sum(p.numel() for p in self.module.parameters()) + 3) // 4 * 4

    def rebuild(self, into: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> None:
        """``into``: (parameter slice, gradient slice) of a :class:`FlatArena` to live in instead of private buffers."""
        params = dict(self.module.named_parameters())
        first = next(iter(params.values()))
        dev = first.device
        order = self._order(params)
        total = sum(params[n].numel() for n in order)
        total_pad = (total + 3) // 4 * 4
        if into is None and self.arena is not None:
            raise RuntimeError('%s lives in a shared FlatArena (one all-reduce bucket for a train step) and one of its parameters was moved '
                               'out of it — rebuild the train step' % type(self.module).__name__)
        if into is not None:
            flat, gflat = into
            assert flat.numel() == total_pad and gflat.numel() == total_pad and flat.device == dev
        else:
            flat = torch.zeros(total_pad, dtype=torch.float32, device=dev)
            gflat = torch.zeros(total_pad, dtype=torch.float32, device=dev)
        off = 0
        self.offsets, self.views, self.gviews = {}, {}, {}
        for n in order:
            p = params[n]
            k = p.numel()
            # keep 16-byte alignment of every tensor start (float4 / TMA friendly)
            off = (off + 3) // 4 * 4 if False else off
            flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = flat[off:off + k].view(p.shape)
            if p.grad is not None:
                gflat[off:off + k].copy_(p.grad.reshape(-1))
                p.grad = gflat[off:off + k].view(p.shape)
            self.offsets[n] = (off, k)
            self.views[n] = p.data
            self.gviews[n] = gflat[off:off + k].view(p.shape)
            off += k
        for fname, names, shape in self.fused:
            o0 = self.offsets[names[0]][0]
            k = sum(self.offsets[n][1] for n in names)
            self.views[fname] = flat[o0:o0 + k].view(*shape)
            self.gviews[fname] = gflat[o0:o0 + k].view(*shape)
        self.flat, self.gflat, self.n = flat, gflat, total
        self._sentinels = [(order[0], params[order[0]]), (order[-1], params[order[-1]])] if order else []
        # BatchNorm counters -> one flat int64 buffer
        owners = [(m, 'num_batches_tracked') for m in self.module.modules()
                  if isinstance(m, nn.modules.batchnorm._BatchNorm) and m.num_batches_tracked is not None]
        if owners:
            nbt = torch.zeros(len(owners), dtype=torch.int64, device=dev)
            for i, (m, key) in enumerate(owners):
                nbt[i] = m._buffers[key].to(dev)
                m._buffers[key] = nbt[i]
            self.nbt = nbt
        # float buffers (BN running stats, TPS constants)
        for name, b in self.module.named_buffers():
            if b.dtype == torch.float32:
                self.views[name] = b

    def ensure(self) -> None:
        if not self.is_current():
            if self.flat is not None and self.frozen:
                raise RuntimeError('the flat parameter buffer of %s was detached (a .to()/.cuda()/.float() after the train step was built moved '
                                   'parameters out of it) — captured CUDA graphs still point at the old buffer; rebuild the train step'
                                   % type(self.module).__name__)
            self.rebuild()
        else:
            for name, b in self.module.named_buffers():
                if b.dtype == torch.float32:
                    self.views[name] = b

    # -- gradient hand-off ------------------------------------------------------------------------------------------
    def live_grads(self) -> set:
        """Names whose ``.grad`` already lives in the flat gradient buffer (so new gradients must accumulate)."""
        live = set()
        for name, p in self.module.named_parameters():
            g = p.grad
            if g is None:
                continue
            gv = self.gviews[name]
            if g.data_ptr() != gv.data_ptr():
                gv.copy_(g)
                p.grad = gv
            live.add(name)
        for fname, names, _ in self.fused:
            if all(n in live for n in names):
                live.add(fname)
            elif any(n in live for n in names):
                raise RuntimeError(f'partial .grad state inside fused parameter group {fname}')
        return live

    def publish_grads(self) -> None:
        """Point every parameter's ``.grad`` at its flat-buffer view (after a backward wrote them)."""
        for name, p in self.module.named_parameters():
            if p.requires_grad:
                p.grad = self.gviews[name]


class FlatArena:
    """ONE parameter buffer and ONE gradient buffer for several modules (the SR model first, then the students): the whole train step
    then needs a single ``ncclAllReduce`` and a single clip+Adam launch pair ("one flat gradient buffer"); the clip
    covers ``[0, n_clipped)`` = the first module."""

    def __init__(self, modules: Sequence[nn.Module]):
        states = [m._flat for m in modules]
        sizes = [st.padded_numel() for st in states]
        dev = next(modules[0].parameters()).device
        self.flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
        self.gflat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
        self.ranges = []
        off = 0
        for st, n in zip(states, sizes):
            st.arena = None
            st.rebuild(into=(self.flat[off:off + n], self.gflat[off:off + n]))
            st.arena = self
            self.ranges.append((off, n))
            off += n
        self.n_clipped = states[0].n
