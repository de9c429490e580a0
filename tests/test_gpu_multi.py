"""2-GPU NCCL parity (SURVEY §8e oracle): two ranks run the PRODUCT train steps (tpgsr_b200.train) on their batch shards with one
ncclAllReduce of the flat gradient arena; afterwards (a) both ranks' flat parameter buffers are bit-identical, (b) they equal the
single-process emulation on the CPU oracle: per-shard gradients (per-rank BatchNorm statistics, per-rank prior dropout) averaged, one
clip, one Adam step.  Needs >= 2 GPUs (`gpurun --gpus 2`); skipped otherwise."""
import json
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu
KW = dict(scale_factor=2, width=128, height=32, STN=True, mask=True, srb_nums=5, hidden_units=32)
PER_RANK = 4


def _smooth(n, seed):
    from oracle.weights import synth_images
    lr, hr = synth_images(n, seed=seed, smooth=True)
    lr[:, 3] = lr[:, :3].mean(1)
    hr[:, 3] = hr[:, :3].mean(1)
    return lr, hr


def _state(shapes):
    from oracle.ops import tps_buffers
    from oracle.weights import synth_state_dict
    sd2 = synth_state_dict(shapes['tsrn_train'], 1)
    sd3 = synth_state_dict(shapes['tsrn_tl_train'], 2)
    for k, v in tps_buffers(16, 64).items():
        sd2['tps.' + k] = v
        sd3['tps.' + k] = v
    return sd2, sd3, synth_state_dict(shapes['crnn'], 7), synth_state_dict(shapes['crnn'], 17)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    from tpgsr_b200 import dist as tdist
    from tpgsr_b200.model.crnn import CRNN
    from tpgsr_b200.model.tsrn import TSRN, TSRN_TL
    from tpgsr_b200.train import TPGSRTrainStep, TSRNTrainStep
    r, local, w = tdist.init_from_env('nccl')
    dev = torch.device('cuda', local)
    shapes = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'shapes.json')))
    sd2, sd3, stu_sd, tea_sd = _state(shapes)
    lr, hr = _smooth(PER_RANK * world, 77)
    lo, hi = tdist.shard_range(PER_RANK * world, r, w)
    lr, hr = lr[lo:hi].to(dev), hr[lo:hi].to(dev)
    # config 2: TSRN
    net = TSRN(**KW)
    net.load_state_dict(sd2, strict=False)
    net = net.to(dev).train()
    ts = TSRNTrainStep(net, batch=PER_RANK, use_graph=True)
    assert ts.world == world
    ts(lr, hr)
    out = {'c2_flat': net._flat.flat.cpu().clone(), 'c2_params': {k: v.detach().cpu().clone() for k, v in net.named_parameters()},
           'c2_norm': ts.opt.grad_norm}
    # config 3: cascade, SR + student in one arena -> one all-reduce
    sr = TSRN_TL(**KW)
    sr.load_state_dict(sd3, strict=False)
    stu, tea = CRNN(32, 1, 37, 256), CRNN(32, 1, 37, 256)
    stu.load_state_dict(stu_sd)
    tea.load_state_dict(tea_sd)
    sr, stu, tea = sr.to(dev).train(), stu.to(dev).train(), tea.to(dev).eval()
    tp = TPGSRTrainStep(sr, [stu], tea, batch=PER_RANK, stu_iter=1, use_distill=True, use_graph=True)
    tp(lr, hr)
    out.update({'c3_flat': tp.arena.flat.cpu().clone(), 'c3_sr': {k: v.detach().cpu().clone() for k, v in sr.named_parameters()},
                'c3_stu': {k: v.detach().cpu().clone() for k, v in stu.named_parameters()}, 'c3_norm': tp.opt.grad_norm})
    torch.cuda.synchronize(dev)
    torch.save(out, os.path.join(out_dir, f'rank{rank}.pt'))
    torch.distributed.destroy_process_group()


def _check_updates(ours, ref, p0, grads, tag, skip=(), frac=0.02):
    worst = []
    for k, r in ref.items():
        if float(grads[k].norm()) < 2e-3 or any(s in k for s in skip):
            continue
        du, dr = (ours[k] - p0[k]).double(), (r - p0[k]).double()
        worst.append((float(((du - dr).abs() > 5e-5).double().mean()), k))       # 5% of lr
    worst.sort(reverse=True)
    assert worst[0][0] < frac, (tag, worst[:6])


@pytest.mark.timeout(900)
def test_two_gpu_steps_bit_identical_and_equal_emulation(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (gpurun --gpus 2)')
    world = 2
    port = 29600 + (os.getpid() % 300)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    a = torch.load(os.path.join(tmp_path, 'rank0.pt'))
    b = torch.load(os.path.join(tmp_path, 'rank1.pt'))
    assert torch.equal(a['c2_flat'], b['c2_flat']) and a['c2_norm'] == b['c2_norm']
    assert torch.equal(a['c3_flat'], b['c3_flat']) and a['c3_norm'] == b['c3_norm']
    # single-process emulation on the CPU oracle
    sys.path.insert(0, ROOT)
    from oracle import ops as o, step as ostep
    shapes = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'shapes.json')))
    sd2, sd3, stu_sd, tea_sd = _state(shapes)
    lr, hr = _smooth(PER_RANK * world, 77)
    # ---- config 2
    per = [ostep.tsrn_train_step({k: v.clone() for k, v in sd2.items()}, lr[i * PER_RANK:(i + 1) * PER_RANK], hr[i * PER_RANK:(i + 1) * PER_RANK], None)
           for i in range(world)]
    keys = list(per[0]['grads'])
    grads = {k: sum(p['grads'][k] for p in per) / world for k in keys}
    raw = {k: g.clone() for k, g in grads.items()}
    norm = o.clip_grad_norm(list(grads.values()), 0.25)
    p = {k: sd2[k].clone() for k in keys}
    with torch.no_grad():
        ostep.AdamState().apply(p, grads)
    assert abs(a['c2_norm'] - float(norm)) < 5e-3 * float(norm), (a['c2_norm'], float(norm))
    _check_updates(a['c2_params'], p, sd2, raw, 'config 2')
    # ---- config 3 (SR clipped, student not; one Adam over both)
    per = [ostep.tpgsr_train_step({k: v.clone() for k, v in sd3.items()}, [dict(stu_sd)], dict(tea_sd), lr[i * PER_RANK:(i + 1) * PER_RANK],
                                  hr[i * PER_RANK:(i + 1) * PER_RANK], None, stu_iter=1) for i in range(world)]
    g_sr = {k: sum(q['sr_grads'][k] for q in per) / world for k in per[0]['sr_grads']}
    g_st = {k: sum(q['stu_grads'][0][k] for q in per) / world for k in per[0]['stu_grads'][0]}
    raw_sr, raw_st = {k: g.clone() for k, g in g_sr.items()}, {k: g.clone() for k, g in g_st.items()}
    norm = o.clip_grad_norm(list(g_sr.values()), 0.25)
    p_sr, p_st = {k: sd3[k].clone() for k in g_sr}, {k: stu_sd[k].clone() for k in g_st}
    with torch.no_grad():
        opt = ostep.AdamState()
        opt.apply({**{'sr.' + k: v for k, v in p_sr.items()}, **{'stu.' + k: v for k, v in p_st.items()}},
                  {**{'sr.' + k: v for k, v in g_sr.items()}, **{'stu.' + k: v for k, v in g_st.items()}})
    assert abs(a['c3_norm'] - float(norm)) < 1e-2 * float(norm), (a['c3_norm'], float(norm))
    _check_updates(a['c3_sr'], p_sr, sd3, raw_sr, 'config 3 SR')
    # the student's conv biases without a BatchNorm behind them carry small gradients: up to 6% of their first Adam steps flip sign
    # (measured; same tensors and size as in tests/test_gpu_parity_full.py::test_config5_srresnet_cascade_step_vs_oracle)
    _check_updates(a['c3_stu'], p_st, stu_sd, raw_st, 'config 3 student', frac=0.10)
