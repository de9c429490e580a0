"""Whole-step execution of the reference's training glue as CUDA graphs.

``TSRNTrainStep`` restates ``interfaces/super_resolution.py:409-424`` for a single SR model (BASELINE config 2):
forward, ``ImageLoss * 100``, backward, ``clip_grad_norm_(0.25)``, ``Adam`` — every kernel is ours, enqueued through
the C ABI, captured once and replayed.  Multi-GPU: one process per GPU, batch sharded by the caller, ONE all-reduce
(sum, fp32) over the flat gradient buffer between the backward graph and the optimiser graph.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import engine, ops
from .optim import FlatAdam


class TSRNTrainStep:
    def __init__(self, model, batch: int, gradient: bool = True, loss_weight=(1.0, 1e-4), lr: float = 1e-3,
                 beta1: float = 0.5, max_norm: float = 0.25, use_graph: bool = True, process_group=None):
        self.model = model
        p0 = next(model.parameters())
        if not p0.is_cuda:
            raise RuntimeError('tpgsr_b200 runs on CUDA only — move the model to a B200 first')
        dev = p0.device
        self.dev = dev
        st = model._flat
        st.ensure()
        self.st = st
        self.gradient, self.w0, self.w1 = bool(gradient), float(loss_weight[0]), float(loss_weight[1])
        self.opt = FlatAdam(st.flat, n_clipped=st.n, lr=lr, betas=(beta1, 0.999), max_norm=max_norm, grads=st.gflat)
        H, W = model.tps_inputsize
        self.lr_buf = torch.zeros(batch, model.in_planes, H, W, device=dev)
        self.hr_buf = torch.zeros(batch, model.in_planes, 2 * H, 2 * W, device=dev)
        self.loss = torch.zeros(1, device=dev)
        self.sr = None
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if (process_group is not None or
                                                                         (torch.distributed.is_available() and torch.distributed.is_initialized())) else 1
        self.use_graph = use_graph
        self._g_fb = None
        self._g_opt = None
        ops.workspace(256 << 20, dev)

    # ---- step pieces ------------------------------------------------------------------------------------------------
    def _fwd_bwd(self):
        m, st = self.model, self.st
        ctx = engine.Ctx(st.views, st.gviews, True)
        if st.nbt is not None:
            ops.add_i64(st.nbt, 1)
        sr, saved = engine.tsrn_fwd(ctx, self.lr_buf, m.stn, m.srb_nums, None)
        ops.image_loss_fwd(sr, self.hr_buf, self.gradient, self.w0, self.w1, 100.0, loss=self.loss, accumulate=False)
        dsr = ops.image_loss_bwd(sr, self.hr_buf, self.gradient, self.w0, self.w1, 100.0)
        engine.tsrn_bwd(ctx, saved, dsr)
        self.sr = sr

    def _optim(self):
        self.opt.step(1.0 / self.world)

    def _capture(self):
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            # warm-up outside capture (also primes the workspace and the caching allocator)
            snap = (self.st.flat.clone(), self.opt.exp_avg.clone(), self.opt.exp_avg_sq.clone(), self.opt.state.clone(),
                    {k: v.clone() for k, v in self.st.views.items() if 'running_' in k},
                    self.st.nbt.clone() if self.st.nbt is not None else None)
            self._fwd_bwd()
            self._optim()
            # undo the warm-up step so capture does not change the trajectory
            self.st.flat.copy_(snap[0]); self.opt.exp_avg.copy_(snap[1]); self.opt.exp_avg_sq.copy_(snap[2])
            self.opt.state.copy_(snap[3])
            for k, v in snap[4].items():
                self.st.views[k].copy_(v)
            if snap[5] is not None:
                self.st.nbt.copy_(snap[5])
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        self._g_fb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._g_fb):
            self._fwd_bwd()
        self._g_opt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._g_opt, pool=self._g_fb.pool()):
            self._optim()

    # ---- public -----------------------------------------------------------------------------------------------------
    def load_batch(self, images_lr: torch.Tensor, images_hr: torch.Tensor) -> None:
        """Copy one batch (host pinned or device) into the step's static input buffers (async on the current stream)."""
        self.lr_buf.copy_(images_lr, non_blocking=True)
        self.hr_buf.copy_(images_hr, non_blocking=True)

    def run(self) -> torch.Tensor:
        """One train step on the currently loaded batch; returns the device scalar loss (ImageLoss*100)."""
        if self.use_graph:
            if self._g_fb is None:
                self._capture()
            self._g_fb.replay()
            if self.world > 1:
                torch.distributed.all_reduce(self.st.gflat, group=self.pg)
            self._g_opt.replay()
        else:
            self._fwd_bwd()
            if self.world > 1:
                torch.distributed.all_reduce(self.st.gflat, group=self.pg)
            self._optim()
        return self.loss

    def __call__(self, images_lr: torch.Tensor, images_hr: torch.Tensor) -> torch.Tensor:
        self.load_batch(images_lr, images_hr)
        return self.run()
