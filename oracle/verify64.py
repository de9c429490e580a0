"""float64 cross-check of the oracle against the reference's own modules (TEST INFRASTRUCTURE; build container only).

    python -m oracle.verify64

In float64 the restatement and the reference agree to ~1e-13 (outputs, input grads, every parameter grad),
which shows the fp32 differences seen in tests/test_oracle_golden.py (1e-4 .. 3e-3 on gradients that pass
through ReLU / max-pool / batch-3 BatchNorm) are round-off amplification, not a modelling difference.
Last run (this container): crnn logits 2.4e-15, crnn input-grad 4.2e-15, tsrn_tl sr 1.2e-13, worst param grad 7.1e-13.
"""
import sys, json, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.gen_golden import _import_reference, load_synth
from oracle import models, ops
from oracle.weights import synth_state_dict, synth_images
m=_import_reference()
torch.set_default_dtype(torch.float64)
net=m['model.crnn.crnn'].CRNN(32,1,37,256).double()
shapes=load_synth(net,7)
net.train()
g=torch.from_numpy(np.random.RandomState(31).rand(3,1,32,100)).requires_grad_(True)
lg=net(g); w=torch.from_numpy(np.random.RandomState(32).randn(*lg.shape)); (lg*w).sum().backward()
p={k:(v.double() if v.is_floating_point() else v) for k,v in synth_state_dict(shapes,7).items()}
g2=g.detach().clone().requires_grad_(True)
for k in p:
    if p[k].is_floating_point() and "running" not in k and "num_batches" not in k: p[k].requires_grad_(True)
lg2=models.crnn(g2,p,True); (lg2*w).sum().backward()
print('crnn logits', float((lg-lg2).norm()/lg.norm()), 'grad', float((g.grad-g2.grad).norm()/g.grad.norm()))
for k,q in net.named_parameters():
    e=float((q.grad-p[k].grad).norm()/(q.grad.norm()+1e-30))
    if e>1e-8 and q.grad.norm()>1e-6: print(k,e)
# TSRN_TL train
kw = dict(scale_factor=2, width=128, height=32, STN=True, mask=True, srb_nums=5, hidden_units=32)
net=m['model.tsrn'].TSRN_TL(**kw).double(); shapes=load_synth(net,2); net.train()
lr,hr=synth_images(4,21); lr=lr.double(); hr=hr.double()
pr=torch.softmax(torch.from_numpy(np.random.RandomState(102).randn(4,37,1,26))*3,1); pr[:1]=0; pr.requires_grad_(True)
sr=net(lr,pr); loss=m['loss.image_loss'].ImageLoss(True,[1,1e-4])(sr,hr)*100; loss.backward()
p={k:(v.double() if v.is_floating_point() else v) for k,v in synth_state_dict(shapes,2).items()}
for k,v in net.tps.state_dict().items(): p['tps.'+k]=v.double()
for k in p:
    if p[k].is_floating_point() and not k.startswith('tps.') and 'running' not in k: p[k].requires_grad_(True)
pr2=pr.detach().clone().requires_grad_(True)
sr2=models.tsrn(lr,p,True,stn=True,text_emb=pr2,tl=True); loss2=ops.image_loss(sr2,hr)*100; loss2.backward()
print('tsrn_tl sr', float((sr-sr2).norm()/sr.norm()), float(loss), float(loss2), 'gprior', float((pr.grad-pr2.grad).norm()/pr.grad.norm()))
worst=0
for k,q in net.named_parameters():
    gg=p[k].grad if p[k].grad is not None else torch.zeros_like(q)
    e=float((q.grad-gg).norm()/(q.grad.norm()+1e-30))
    if q.grad.norm()>1e-9: worst=max(worst,e)
print('worst param grad rel', worst)
