"""world_size=2 gloo test (CPU) of the data-parallel step semantics: shard the batch, per-rank BatchNorm,
ONE sum all-reduce of the flat gradient buffer, then clip + Adam with grad_scale = 1/world on every rank.

The arithmetic on each rank is the CPU oracle (the CUDA kernels need a GPU); what is under test is the host-side
protocol in tpgsr_b200.dist and the claim that replicas stay bit-identical and equal the single-process emulation."""
import json
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _flat(named):
    keys = sorted(named)
    return keys, torch.cat([named[k].reshape(-1) for k in keys])


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.set_num_threads(2)
    from oracle import ops as o, step
    from oracle.ops import tps_buffers
    from oracle.weights import synth_images, synth_state_dict
    from tpgsr_b200 import dist as tdist
    r, _, w = tdist.init_from_env('gloo')
    assert (r, w) == (rank, world)
    shapes = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'shapes.json')))['tsrn_train']
    p = synth_state_dict(shapes, 1)
    p.update({'tps.' + k: v for k, v in tps_buffers(16, 64).items()})
    lr, hr = synth_images(4, seed=77, smooth=True)
    lo, hi = tdist.shard_range(4, rank, world)
    out = step.tsrn_train_step(p, lr[lo:hi], hr[lo:hi], None)          # per-rank BN statistics, local mean loss
    keys, g = _flat(out['grads'])
    tdist.all_reduce_flat_(g)                                            # the one collective of the step
    g = g / world
    off = 0
    grads = {}
    for k in keys:
        n = p[k].numel()
        grads[k] = g[off:off + n].reshape(p[k].shape).clone()
        off += n
    norm = o.clip_grad_norm(list(grads.values()), 0.25)
    opt = step.AdamState()
    with torch.no_grad():
        opt.apply({k: p[k] for k in keys}, grads)
    torch.save({'params': {k: p[k] for k in keys}, 'norm': norm, 'grads_raw': out['grads']}, os.path.join(out_dir, f'rank{rank}.pt'))
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_step_matches_single_process_emulation(tmp_path):
    port = 29500 + (os.getpid() % 500)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = torch.load(os.path.join(tmp_path, 'rank0.pt'))
    b = torch.load(os.path.join(tmp_path, 'rank1.pt'))
    # replicas are bit-identical after the step
    for k in a['params']:
        assert torch.equal(a['params'][k], b['params'][k]), k
    assert float(a['norm']) == float(b['norm'])
    # and equal to the single-process emulation: average of the per-shard gradients, one clip, one Adam step
    sys.path.insert(0, ROOT)
    from oracle import ops as o, step
    from oracle.ops import tps_buffers
    from oracle.weights import synth_state_dict
    shapes = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'shapes.json')))['tsrn_train']
    p = synth_state_dict(shapes, 1)
    p.update({'tps.' + k: v for k, v in tps_buffers(16, 64).items()})
    keys = sorted(a['params'])
    grads = {k: (a['grads_raw'][k] + b['grads_raw'][k]) / 2 for k in keys}
    norm = o.clip_grad_norm(list(grads.values()), 0.25)
    opt = step.AdamState()
    with torch.no_grad():
        opt.apply({k: p[k] for k in keys}, grads)
    assert abs(float(norm) - float(a['norm'])) < 1e-6 * float(norm)
    for k in keys:
        assert torch.allclose(p[k], a['params'][k], rtol=0, atol=1e-7), k


def test_shard_range():
    sys.path.insert(0, ROOT)
    from tpgsr_b200.dist import shard_range
    assert [shard_range(384, r, 8) for r in (0, 7)] == [(0, 48), (336, 384)]
    with pytest.raises(ValueError):
        shard_range(50, 0, 8)
