"""Oracle train / inference steps (TEST INFRASTRUCTURE — see ``oracle/__init__.py``).

Restates the step glue of ``interfaces/super_resolution.py:297-424`` on top of the functional
oracle models: teacher prior, student cascade, ImageLoss*100 (+ SemanticLoss*100), backward,
``clip_grad_norm_(SR params, 0.25)``, ``Adam(lr 1e-3, betas (0.5, 0.999))``.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import models, ops

Tensor = torch.Tensor
P = Dict[str, Tensor]

_NON_PARAM_LEAVES = ('running_mean', 'running_var', 'num_batches_tracked')


def is_param(name: str) -> bool:
    return not (name.rsplit('.', 1)[-1] in _NON_PARAM_LEAVES or name.startswith('tps.'))


class AdamState:
    """Per-parameter m/v/step for the oracle's Adam (interfaces/base.py:449-450)."""

    def __init__(self, lr=1e-3, beta1=0.5, beta2=0.999, eps=1e-8):
        self.lr, self.b1, self.b2, self.eps = lr, beta1, beta2, eps
        self.m: Dict[str, Tensor] = {}
        self.v: Dict[str, Tensor] = {}
        self.t = 0

    def apply(self, named_params: Dict[str, Tensor], named_grads: Dict[str, Tensor]):
        self.t += 1
        for k, p in named_params.items():
            g = named_grads.get(k)
            if g is None:
                continue
            if k not in self.m:
                self.m[k] = torch.zeros_like(p)
                self.v[k] = torch.zeros_like(p)
            ops.adam_step(p, g, self.m[k], self.v[k], self.t, self.lr, self.b1, self.b2, self.eps)


def _leaf_params(p: P) -> Dict[str, Tensor]:
    out = {}
    for k, v in p.items():
        if is_param(k):
            p[k] = v.detach().requires_grad_(True)
            out[k] = p[k]
    return out


def tsrn_train_step(p: P, lr_img: Tensor, hr_img: Tensor, opt: Optional[AdamState] = None,
                    stn: bool = True, gradient: bool = True, srb_nums: int = 5):
    """Config 2: TSRN(STN, mask) + ImageLoss(gradient,[1,1e-4]); loss*100; clip 0.25; Adam.
    Returns dict(loss, sr, grads)."""
    leaves = _leaf_params(p)
    sr = models.tsrn(lr_img, p, training=True, stn=stn, srb_nums=srb_nums)
    loss = ops.image_loss(sr, hr_img, gradient) * 100
    grads = torch.autograd.grad(loss, list(leaves.values()), allow_unused=True)
    named = {k: (g if g is not None else torch.zeros_like(leaves[k])) for k, g in zip(leaves, grads)}
    raw = {k: g.clone() for k, g in named.items()}
    out = {'loss': loss.detach(), 'sr': sr.detach(), 'grads': raw}
    if opt is not None:
        out['grad_norm'] = ops.clip_grad_norm(list(named.values()), 0.25)
        with torch.no_grad():
            for k in leaves:
                p[k] = p[k].detach()
            opt.apply({k: p[k] for k in leaves}, named)
    else:
        for k in leaves:
            p[k] = p[k].detach()
    return out


def _sr_model(arch: str, lr_img: Tensor, sr_p: P, training: bool, stn: bool, prior: Tensor) -> Tensor:
    if arch == 'srresnet_tl':
        return models.srresnet_tl(lr_img, sr_p, training=training, text_emb=prior)
    return models.tsrn(lr_img, sr_p, training=training, stn=stn, text_emb=prior, tl=True)


def tpgsr_train_step(sr_p: P, students: List[P], teacher: P, lr_img: Tensor, hr_img: Tensor,
                     opt: Optional[AdamState] = None, stu_iter: int = 1, use_distill: bool = True,
                     stn: bool = True, gradient: bool = True, tpg_share: bool = False, arch: str = 'tsrn_tl'):
    """Configs 3/4: interfaces/super_resolution.py:297-424 with ``sr_share`` (one SR model).
    Students are in train mode, teacher in eval mode; Adam covers SR + student params, the clip only SR.
    Config 5: ``arch='srresnet_tl'``, ``gradient=False`` (nn.MSELoss, interfaces/base.py:359-362)."""
    sr_leaves = _leaf_params(sr_p)
    stu_leaves = [_leaf_params(s) for s in students]
    with torch.no_grad():
        hr_gray = ops.parse_crnn_data(hr_img)
        label_hr = torch.softmax(models.crnn(hr_gray, teacher, training=False), -1)
    cascade = lr_img
    loss_img = 0.
    loss_dis = 0.
    for i in range(stu_iter):
        sp = students[0 if tpg_share else i]
        gray = ops.parse_crnn_data(cascade)
        logits = models.crnn(gray, sp, training=True)
        probs, prior = models.prior_from_logits(logits, True)
        if use_distill:
            loss_dis = loss_dis + ops.semantic_loss(probs, label_hr) * 100
        cascade = _sr_model(arch, lr_img, sr_p, True, stn, prior)
        loss_img = loss_img + ops.image_loss(cascade, hr_img, gradient) * 100
    loss = loss_img + loss_dis
    all_leaves = list(sr_leaves.values()) + [t for sl in stu_leaves for t in sl.values()]
    grads = torch.autograd.grad(loss, all_leaves, allow_unused=True)
    grads = [g if g is not None else torch.zeros_like(t) for g, t in zip(grads, all_leaves)]
    n_sr = len(sr_leaves)
    sr_g = dict(zip(sr_leaves, grads[:n_sr]))
    stu_g = []
    off = n_sr
    for sl in stu_leaves:
        stu_g.append(dict(zip(sl, grads[off:off + len(sl)])))
        off += len(sl)
    out = {'loss': loss.detach(), 'loss_img': torch.as_tensor(loss_img).detach(),
           'loss_distill': torch.as_tensor(loss_dis).detach(), 'sr': cascade.detach(),
           'sr_grads': {k: g.clone() for k, g in sr_g.items()},
           'stu_grads': [{k: g.clone() for k, g in d.items()} for d in stu_g]}
    for d, leaves in [(sr_p, sr_leaves)] + list(zip(students, stu_leaves)):
        for k in leaves:
            d[k] = d[k].detach()
    if opt is not None:
        out['grad_norm'] = ops.clip_grad_norm(list(sr_g.values()), 0.25)
        with torch.no_grad():
            named_p = {('sr.' + k): sr_p[k] for k in sr_leaves}
            named_g = {('sr.' + k): g for k, g in sr_g.items()}
            for j, (s, sl) in enumerate(zip(students, stu_leaves)):
                named_p.update({(f'stu{j}.' + k): s[k] for k in sl})
                named_g.update({(f'stu{j}.' + k): g for k, g in stu_g[j].items()})
            opt.apply(named_p, named_g)
    return out


@torch.no_grad()
def tpgsr_infer(sr_p: P, students: List[P], lr_img: Tensor, stu_iter: int = 1, tpg_share: bool = False, arch: str = 'tsrn_tl') -> Tensor:
    """Evaluation data-flow (interfaces/super_resolution.py:748-768): eval mode, no STN, no prior dropout."""
    cascade = lr_img
    for i in range(stu_iter):
        sp = students[0 if tpg_share else i]
        logits = models.crnn(ops.parse_crnn_data(cascade), sp, training=False)
        _, prior = models.prior_from_logits(logits, False)
        cascade = _sr_model(arch, lr_img, sr_p, False, True, prior)
    return cascade
