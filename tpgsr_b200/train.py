"""Whole-step execution of the reference's training glue as CUDA graphs.

``TSRNTrainStep`` restates ``interfaces/super_resolution.py:409-424`` for a single SR model (BASELINE config 2):
forward, ``ImageLoss * 100``, backward, ``clip_grad_norm_(0.25)``, ``Adam`` — every kernel is ours, enqueued through
the C ABI, captured once and replayed.  Multi-GPU: one process per GPU, batch sharded by the caller, ONE all-reduce
(sum, fp32) over the flat gradient buffer between the backward graph and the optimiser graph.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import dist as tdist, engine, ops
from .model._flat import FlatArena
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
        # packed weight operands persist across steps and are refreshed by one launch after Adam (ops.PackStore)
        self.packs = ops.PackStore() if ops.PREPACK else None
        ops.prime_workspaces(256 << 20, dev)

    # ---- step pieces ------------------------------------------------------------------------------------------------
    def refresh_packs(self) -> None:
        """Call after changing the model's parameters from outside the step (load_state_dict, manual edits)."""
        if self.packs is not None:
            self.packs.refresh()

    def _fwd_bwd(self):
        m, st = self.model, self.st
        ctx = engine.Ctx(st.views, st.gviews, True, store=self.packs)
        if st.nbt is not None:
            ops.add_i64(st.nbt, 1)
        sr, saved = engine.tsrn_fwd(ctx, self.lr_buf, m.stn, m.srb_nums, None)
        ops.image_loss_fwd(sr, self.hr_buf, self.gradient, self.w0, self.w1, 100.0, loss=self.loss, accumulate=False)
        dsr = ops.image_loss_bwd(sr, self.hr_buf, self.gradient, self.w0, self.w1, 100.0)
        engine.tsrn_bwd(ctx, saved, dsr)
        ctx.join()                      # side-stream weight gradients are complete before the graph / the all-reduce continues
        self.sr = sr

    def _optim(self):
        self.opt.step(1.0 / self.world)
        self.refresh_packs()

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
            self.refresh_packs()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        self._g_fb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._g_fb):
            self._fwd_bwd()
        self._g_opt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._g_opt, pool=self._g_fb.pool()):
            self._optim()
        self.st.frozen = True            # the graphs hold the flat buffers' addresses: a later rebuild must raise, not silently detach

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
            self.st.assert_attached()
            self._g_fb.replay()
            tdist.all_reduce_flat_(self.st.gflat, self.pg)
            self._g_opt.replay()
        else:
            self._fwd_bwd()
            tdist.all_reduce_flat_(self.st.gflat, self.pg)
            self._optim()
        return self.loss

    def __call__(self, images_lr: torch.Tensor, images_hr: torch.Tensor) -> torch.Tensor:
        self.load_batch(images_lr, images_hr)
        return self.run()


class TPGSRTrainStep:
    """The cascade step of ``interfaces/super_resolution.py:297-424`` (arch ``tsrn_tl_cascade``, ``--sr_share``) as CUDA graphs:

        teacher prior  : CRNN(eval)(parse_crnn_data(HR)) -> softmax                      (no grad)
        for i < stu_iter: gray = parse_crnn_data(cascade); logits = student_i(gray); probs, prior = softmax / permute / dropout
                          loss += 100 * SemanticLoss(probs, teacher) [use_distill]; cascade = TSRN_TL(LR, prior)
                          loss += 100 * ImageLoss(cascade, HR)
        backward through every stage (the stage-i SR output feeds stage i+1's recogniser un-detached),
        clip_grad_norm_(SR params, 0.25), Adam over SR + student parameters.

    Configs 3 (stu_iter=1) and 4 (stu_iter=3) of BASELINE.json, and config 5 when ``model`` is an ``SRResNet_TL`` (arch ``srresnet_tl`` shares
    the loop, ``ABLATION_SET`` at interfaces/super_resolution.py:44; its criterion is ``nn.MSELoss`` = ``gradient=False``,
    interfaces/base.py:359-362).  ``students`` may hold one module (``tpg_share``) or stu_iter."""

    def __init__(self, model, students, teacher, batch: int, stu_iter: int = 1, use_distill: bool = True, gradient: bool = True,
                 loss_weight=(1.0, 1e-4), lr: float = 1e-3, beta1: float = 0.5, max_norm: float = 0.25, use_graph: bool = True,
                 process_group=None):
        from . import crnn_engine, infogen, ops_prior  # noqa: F401
        self.model, self.students, self.teacher = model, list(students), teacher
        p0 = next(model.parameters())
        if not p0.is_cuda:
            raise RuntimeError('tpgsr_b200 runs on CUDA only — move the models to a B200 first')
        self.dev = p0.device
        self.stu_iter, self.use_distill = stu_iter, use_distill
        self._srresnet = type(model).__name__ == 'SRResNet_TL'      # config 5: pass gradient=False (the reference pairs it with nn.MSELoss)
        self.gradient, self.w0, self.w1 = bool(gradient), float(loss_weight[0]), float(loss_weight[1])
        self.tpg_share = len(self.students) == 1 and stu_iter > 1
        assert self.tpg_share or len(self.students) == stu_iter
        teacher._flat.ensure()
        # SR model + students share one parameter / gradient arena: ONE all-reduce and ONE clip+Adam launch pair per step; the clip
        # covers the SR model's range only (interfaces/super_resolution.py:422-423), Adam everything (interfaces/base.py:449-450)
        self.arena = FlatArena([model] + self.students)
        self.opt = FlatAdam(self.arena.flat, n_clipped=self.arena.n_clipped, lr=lr, betas=(beta1, 0.999), max_norm=max_norm,
                            grads=self.arena.gflat)
        self.opt_sr = self.opt           # (grad_norm of the clipped range = the SR model's, as clip_grad_norm_ returns it)
        H, W = model.tps_inputsize
        self.lr_buf = torch.zeros(batch, model.in_planes, H, W, device=self.dev)
        self.hr_buf = torch.zeros(batch, model.in_planes, 2 * H, 2 * W, device=self.dev)
        self.batch = batch
        self.loss = torch.zeros(1, device=self.dev)
        self.loss_img = torch.zeros(1, device=self.dev)
        self.loss_distill = torch.zeros(1, device=self.dev)
        self.sr = None
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if (torch.distributed.is_available() and torch.distributed.is_initialized()) else 1
        self.use_graph = use_graph
        self._g_fb = None
        self._g_opt = None
        # packed weight operands persist across steps (ops.PackStore): SR model and students are refreshed by one launch each after Adam,
        # the frozen teacher's are computed once
        # the frozen teacher's packed weights are computed ONCE and kept (a PackStore that the optimiser never refreshes; call
        # refresh_packs() after changing the teacher's parameters from outside)
        self.packs = {'teacher': ops.PackStore()}
        if ops.PREPACK:
            self.packs['sr'] = ops.PackStore()
            self.packs.update({('stu', i): ops.PackStore() for i in range(len(self.students))})
        ops.prime_workspaces(256 << 20, self.dev)

    # ---- SR model adapters: TSRN_TL (configs 3 / 4) or SRResNet_TL (config 5; model/srresnet.py:139-163) -------------------------------
    def _sr_forward(self, ctx, prior_nhwc):
        m = self.model
        if self._srresnet:
            from .model import srresnet as sres
            H, W = self.lr_buf.shape[2], self.lr_buf.shape[3]
            text, s_text = sres.info_gen2d_fwd(ctx, None, (H, W), prior_nhwc=prior_nhwc)
            sr, s_sr = sres.srresnet_fwd(ctx, self.lr_buf, text)
        else:
            from . import infogen
            text, s_text = infogen.info_gen_fwd(ctx, self.lr_buf, None, m.emb_cls, prior_nhwc=prior_nhwc)
            sr, s_sr = engine.tsrn_fwd(ctx, self.lr_buf, m.stn, m.srb_nums, text)
        return sr, s_text, s_sr

    def _sr_backward(self, ctx, s_text, s_sr, dsr):
        dtext = torch.zeros_like(s_text['text'])
        if self._srresnet:
            from .model import srresnet as sres
            sres.srresnet_bwd(ctx, s_sr, dsr, dtext)
            return sres.info_gen2d_bwd(ctx, s_text, dtext, True, nhwc_out=True)
        from . import infogen
        engine.tsrn_bwd(ctx, s_sr, dsr, dtext)
        return infogen.info_gen_bwd(ctx, s_text, dtext, True, nhwc_out=True)

    def _fwd_bwd(self):
        from . import crnn_engine, infogen, ops_prior
        m = self.model
        B = self.batch
        drop = B // 4
        st_sr = m._flat
        # ---- teacher prior (eval mode, no grad).  It depends on the HR batch only, so it runs on the side stream (idle during the forward
        #      pass) concurrently with the first student / SR stage; the main stream joins right before the first distillation loss.
        side = ops.side_stream(self.dev) if self.use_distill else None
        with ops.on_stream_after_current(side, self.dev):
            tctx = engine.Ctx(self.teacher._flat.views, None, False, store=self.packs.get('teacher'))
            gray_hr = ops_prior.bicubic_gray_fwd(self.hr_buf)
            logits_hr, _ = crnn_engine.crnn_fwd(tctx, gray_hr)
            label_hr, _ = ops_prior.softmax_prior_fwd(logits_hr, 0, need_prior=False)
        teacher_joined = side is None
        # ---- cascade forward
        stages = []
        cascade = self.lr_buf
        sr_ctx0 = None
        stu_ctx = {}
        first_img, first_dis = True, True
        for i in range(self.stu_iter):
            si = 0 if self.tpg_share else i
            stu = self.students[si]
            sctx = engine.Ctx(stu._flat.views, stu._flat.gviews, True, share=stu_ctx.get(si), store=self.packs.get(('stu', si)))
            stu_ctx.setdefault(si, sctx)
            if stu._flat.nbt is not None:
                ops.add_i64(stu._flat.nbt, 1)
            gray = ops_prior.bicubic_gray_fwd(cascade)
            logits, s_crnn = crnn_engine.crnn_fwd(sctx, gray, need_input_grad=i > 0)
            probs, prior = ops_prior.softmax_prior_fwd(logits, drop)
            if self.use_distill:
                if not teacher_joined:
                    ops.join_side(self.dev)
                    teacher_joined = True
                ops.semantic_loss_fwd(probs, label_hr, 100.0, loss=self.loss_distill, accumulate=not first_dis)
                first_dis = False
            ctx = engine.Ctx(st_sr.views, st_sr.gviews, True, share=sr_ctx0, store=self.packs.get('sr'))
            sr_ctx0 = sr_ctx0 or ctx
            if st_sr.nbt is not None:
                ops.add_i64(st_sr.nbt, 1)
            sr, s_text, s_sr = self._sr_forward(ctx, prior)
            ops.image_loss_fwd(sr, self.hr_buf, self.gradient, self.w0, self.w1, 100.0, loss=self.loss_img, accumulate=not first_img)
            first_img = False
            stages.append((sctx, s_crnn, probs, ctx, s_text, s_sr, sr))
            cascade = sr
        if not self.use_distill:
            self.loss_distill.zero_()
        ops.add(self.loss_img, self.loss_distill, out=self.loss)
        # ---- cascade backward
        dnext = None                                  # d(loss)/d(SR output of the stage) contributed by the NEXT stage's recogniser
        for i in range(self.stu_iter - 1, -1, -1):
            sctx, s_crnn, probs, ctx, s_text, s_sr, sr = stages[i]
            dsr = ops.image_loss_bwd(sr, self.hr_buf, self.gradient, self.w0, self.w1, 100.0, dout=dnext, accumulate=dnext is not None)
            dprior = self._sr_backward(ctx, s_text, s_sr, dsr)
            dprobs = ops.semantic_loss_bwd(probs, label_hr, 100.0) if self.use_distill else None
            dlogits = ops_prior.softmax_prior_bwd(probs, dprobs, dprior, drop)
            dgray = crnn_engine.crnn_bwd(sctx, s_crnn, dlogits)
            if i > 0:
                dnext = torch.zeros_like(stages[i - 1][6])
                ops_prior.bicubic_gray_bwd(dgray, dnext, False)
        ops.join_side(self.dev)          # side-stream weight gradients are complete before the graph / the all-reduce continues
        self.sr = stages[-1][6]

    def refresh_packs(self, teacher: bool = True) -> None:
        """Call after changing any module's parameters from outside the step (load_state_dict, manual edits)."""
        for k, st in self.packs.items():
            if teacher or k != 'teacher':
                st.refresh()

    def _optim(self):
        self.opt.step(1.0 / self.world)
        self.refresh_packs(teacher=False)        # the teacher is frozen: its packed weights are computed once

    def _all_reduce(self):
        tdist.all_reduce_flat_(self.arena.gflat, self.pg)      # the one collective of the step

    def _capture(self):
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        mods = [self.model] + self.students
        o = self.opt
        with torch.cuda.stream(s):
            snap_opt = (self.arena.flat.clone(), o.exp_avg.clone(), o.exp_avg_sq.clone(), o.state.clone())
            snap = [({k: v.clone() for k, v in m._flat.views.items() if 'running_' in k},
                     m._flat.nbt.clone() if m._flat.nbt is not None else None) for m in mods]
            self._fwd_bwd()
            self._optim()
            # undo the warm-up step so capture does not change the trajectory
            self.arena.flat.copy_(snap_opt[0]); o.exp_avg.copy_(snap_opt[1]); o.exp_avg_sq.copy_(snap_opt[2]); o.state.copy_(snap_opt[3])
            for m, sn in zip(mods, snap):
                for k, v in sn[0].items():
                    m._flat.views[k].copy_(v)
                if sn[1] is not None:
                    m._flat.nbt.copy_(sn[1])
            self.refresh_packs(teacher=False)
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        self._g_fb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._g_fb):
            self._fwd_bwd()
        self._g_opt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._g_opt, pool=self._g_fb.pool()):
            self._optim()
        for m in mods + [self.teacher]:
            m._flat.frozen = True        # the graphs hold the flat buffers' addresses: a later rebuild must raise, not silently detach

    def load_batch(self, images_lr, images_hr):
        self.lr_buf.copy_(images_lr, non_blocking=True)
        self.hr_buf.copy_(images_hr, non_blocking=True)

    def run(self):
        if self.use_graph:
            if self._g_fb is None:
                self._capture()
            for m in [self.model] + self.students:
                m._flat.assert_attached()
            self._g_fb.replay()
            self._all_reduce()
            self._g_opt.replay()
        else:
            self._fwd_bwd()
            self._all_reduce()
            self._optim()
        return self.loss

    def __call__(self, images_lr, images_hr):
        self.load_batch(images_lr, images_hr)
        return self.run()


@torch.no_grad()
def tpgsr_infer(model, students, images_lr: torch.Tensor, stu_iter: int = 1) -> torch.Tensor:
    """Evaluation data-flow of ``interfaces/super_resolution.py:748-768`` (eval mode: no STN, running BN stats, no prior
    dropout) on the CUDA kernels; ``model`` / ``students`` must be in ``.eval()``."""
    from . import crnn_engine, infogen, ops_prior
    model._flat.ensure()
    cascade = images_lr
    for i in range(stu_iter):
        stu = students[0 if len(students) == 1 else i]
        stu._flat.ensure()
        sctx = engine.Ctx(stu._flat.views, None, False)
        logits, _ = crnn_engine.crnn_fwd(sctx, ops_prior.bicubic_gray_fwd(cascade))
        _, prior = ops_prior.softmax_prior_fwd(logits, 0)
        ctx = engine.Ctx(model._flat.views, None, False)
        if type(model).__name__ == 'SRResNet_TL':
            from .model import srresnet as sres
            text, _ = sres.info_gen2d_fwd(ctx, None, (images_lr.shape[2], images_lr.shape[3]), prior_nhwc=prior)
            cascade, _ = sres.srresnet_fwd(ctx, images_lr, text)
        else:
            text, _ = infogen.info_gen_fwd(ctx, images_lr, None, model.emb_cls, prior_nhwc=prior)
            cascade, _ = engine.tsrn_fwd(ctx, images_lr, model.stn, model.srb_nums, text)
    return cascade
