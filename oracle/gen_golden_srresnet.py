"""Golden fixture for BASELINE config 5 (SRResNet_TL + text prior, MSE loss): imports the reference's own ``model/srresnet.py`` (read-only,
``/root/reference``) exactly like ``oracle/gen_golden.py`` does, loads the deterministic synthetic weights, runs one train-mode forward /
backward (``nn.MSELoss()*100``, interfaces/base.py:359-362, interfaces/super_resolution.py:385) and one eval forward, and stores inputs,
outputs and per-parameter gradient fingerprints in ``tests/golden/srresnet_tl.npz`` (+ the parameter shapes in ``shapes.json``).

    python -m oracle.gen_golden_srresnet
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from oracle.gen_golden import OUT, _import_reference, buffers_fp, grads_fp, load_synth, save


def main():
    torch.set_num_threads(8)
    from oracle.weights import synth_images
    m = _import_reference()
    torch.manual_seed(0)
    net = m['model.srresnet'].SRResNet_TL(scale_factor=2, width=128, height=32, STN=False, mask=True)
    shapes = load_synth(net, 9)
    # PReLU slopes: distinct values per block so that a wrong slope / slope gradient shows up
    with torch.no_grad():
        for i, (k, p) in enumerate(sorted((k, p) for k, p in net.named_parameters() if p.numel() == 1)):
            p.fill_(0.1 + 0.05 * i)
    B = 4
    lr_i, hr_i = synth_images(B, seed=31, smooth=True)
    rr = np.random.RandomState(131)
    pr = torch.softmax(torch.from_numpy(rr.randn(B, 37, 1, 26).astype(np.float32)) * 3, 1)
    pr[: B // 4] = 0
    pr.requires_grad_(True)
    net.train()
    sr = net(lr_i, pr)
    loss = torch.nn.MSELoss()(sr, hr_i) * 100
    loss.backward()
    out = {'lr': lr_i, 'hr': hr_i, 'prior': pr, 'sr': sr, 'loss': loss, 'grad_prior': pr.grad}
    out.update(grads_fp(net))
    out.update(buffers_fp(net))
    out.update({'slope/' + k: p.detach().clone() for k, p in net.named_parameters() if p.numel() == 1})
    net.eval()
    with torch.no_grad():
        out['sr_eval'] = net(lr_i, pr.detach())
        out['sr_eval_noprior'] = net(lr_i)
    save('srresnet_tl', **out)
    path = os.path.join(OUT, 'shapes.json')
    allshapes = json.load(open(path))
    allshapes['srresnet_tl'] = {n: list(s) for n, s in shapes.items()}
    with open(path, 'w') as f:
        json.dump(allshapes, f, indent=0)
    print('done', len(shapes), 'tensors')


if __name__ == '__main__':
    main()
