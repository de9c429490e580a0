"""CRNN drop-in module (reference: ``model/crnn/crnn.py:5-90``; constructed as ``CRNN(32, 1, 37, 256)`` at
``interfaces/base.py:635``): same constructor, ``forward(input)`` and 49 ``state_dict`` keys; executed by
``tpgsr_b200.crnn_engine`` on the CUDA library (no PyTorch fallback)."""
from __future__ import annotations

import torch
from torch import nn

from .. import crnn_engine, engine, ops
from ._flat import FlatState


class BidirectionalLSTM(nn.Module):
    def __init__(self, nIn, nHidden, nOut):
        super().__init__()
        self.rnn = nn.LSTM(nIn, nHidden, bidirectional=True)
        self.embedding = nn.Linear(nHidden * 2, nOut)


class _CRNNFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, gray, anchor, want_backward):
        logits, saved = module._run_forward(gray, want_backward, gray.requires_grad)
        ctx.module, ctx.saved = module, saved
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        dgray = ctx.module._run_backward(ctx.saved, dlogits)
        ctx.saved = None
        return None, dgray, None, None


class CRNN(nn.Module):
    def __init__(self, imgH, nc, nclass, nh, n_rnn=2, leakyRelu=False):
        super().__init__()
        assert imgH % 16 == 0, 'imgH has to be a multiple of 16'
        if leakyRelu or nh != 256 or nc != 1:
            raise NotImplementedError('tpgsr_b200 CRNN kernels cover the reference configuration CRNN(32, 1, 37, 256)')
        ks, ps, nm = [3, 3, 3, 3, 3, 3, 2], [1, 1, 1, 1, 1, 1, 0], [64, 128, 256, 256, 512, 512, 512]
        cnn = nn.Sequential()
        pools = {0: nn.MaxPool2d(2, 2), 1: nn.MaxPool2d(2, 2), 3: nn.MaxPool2d((2, 2), (2, 1), (0, 1)),
                 5: nn.MaxPool2d((2, 2), (2, 1), (0, 1))}
        pool_names = {0: 'pooling0', 1: 'pooling1', 3: 'pooling2', 5: 'pooling3'}
        for i in range(7):
            cin = nc if i == 0 else nm[i - 1]
            cnn.add_module(f'conv{i}', nn.Conv2d(cin, nm[i], ks[i], 1, ps[i]))
            if i in (2, 4, 6):
                cnn.add_module(f'batchnorm{i}', nn.BatchNorm2d(nm[i]))
            cnn.add_module(f'relu{i}', nn.ReLU(True))
            if i in pools:
                cnn.add_module(pool_names[i], pools[i])
        self.cnn = cnn
        self.rnn = nn.Sequential(BidirectionalLSTM(512, nh, nh), BidirectionalLSTM(nh, nh, nclass))
        fused = []
        for k, nin in ((0, 512), (1, nh)):
            g = f'rnn.{k}.rnn.'
            fused += [(g + 'weight_ih', [g + 'weight_ih_l0', g + 'weight_ih_l0_reverse'], (8 * nh, nin)),
                      (g + 'weight_hh', [g + 'weight_hh_l0', g + 'weight_hh_l0_reverse'], (2, 4 * nh, nh)),
                      (g + 'bias_ih', [g + 'bias_ih_l0', g + 'bias_ih_l0_reverse'], (8 * nh,)),
                      (g + 'bias_hh', [g + 'bias_hh_l0', g + 'bias_hh_l0_reverse'], (2, 4 * nh))]
        self._flat = FlatState(self, fused)

    def _run_forward(self, gray, want_backward: bool, need_input_grad: bool):
        if not gray.is_cuda:
            raise RuntimeError('tpgsr_b200 modules run on CUDA (sm_100a) only — there is no CPU fallback')
        if gray.dtype != torch.float32 or gray.dim() != 4 or gray.shape[1] != 1:
            raise ValueError(f'expected float32 [B,1,32,W] input, got {tuple(gray.shape)} {gray.dtype}')
        st = self._flat
        st.ensure()
        train_bwd = want_backward and (need_input_grad or any(p.requires_grad for p in self.parameters()))
        ectx = engine.Ctx(st.views, st.gviews if train_bwd else None, self.training)
        if self.training and st.nbt is not None:
            ops.add_i64(st.nbt, 1)
        logits, saved = crnn_engine.crnn_fwd(ectx, gray, need_input_grad and train_bwd)
        return logits, (ectx, saved) if train_bwd else None

    def _run_backward(self, saved_all, dlogits):
        if saved_all is None:
            raise RuntimeError('backward through a CRNN forward that ran without grad enabled')
        ectx, saved = saved_all
        st = self._flat
        ectx._written = set(st.live_grads())
        dgray = crnn_engine.crnn_bwd(ectx, saved, dlogits.contiguous())
        ectx.join()
        members = {f: names for f, names, _ in st.fused}
        params = dict(self.named_parameters())
        for name in ectx._written:
            for n in members.get(name, [name]):
                p = params.get(n)
                if p is not None and p.requires_grad:
                    p.grad = st.gviews[n]
        return dgray

    @property
    def module(self):
        """The un-wrapped model (callers written for ``nn.DataParallel`` read ``.module``; one process per GPU has no wrapper)."""
        return self

    def complexity_info(self, input_res):
        """Multiply-accumulates of one ``(1, 32, W)`` image (what ``ptflops`` prints for the reference module, interfaces/base.py:638)."""
        _, H, W = input_res
        macs, h, w, cin = 0, H, W, 1
        for i, (co, k, p) in enumerate(zip([64, 128, 256, 256, 512, 512, 512], [3, 3, 3, 3, 3, 3, 2], [1, 1, 1, 1, 1, 1, 0])):
            h, w = h + 2 * p - k + 1, w + 2 * p - k + 1
            macs += cin * co * k * k * h * w
            cin = co
            if i in (0, 1):
                h, w = h // 2, w // 2
            elif i in (3, 5):
                h, w = h // 2, w + 1
        T, nh = w, 256
        macs += T * 2 * 4 * nh * (512 + nh) + T * 2 * nh * nh            # BiLSTM 1 + embedding
        macs += T * 2 * 4 * nh * (nh + nh) + T * 2 * nh * self.rnn[1].embedding.out_features
        return macs

    def forward(self, input):
        anchor = next(self.parameters())
        return _CRNNFunction.apply(self, input, anchor, torch.is_grad_enabled())
