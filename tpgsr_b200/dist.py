"""Data-parallel plumbing: one process per GPU, the batch sharded contiguously, ONE all-reduce per step.

Replaces the reference's ``nn.DataParallel`` replicate/scatter/gather/reduce (``interfaces/base.py:390-400``) with the
scheme of SURVEY.md §8e: every rank holds full replicas, BatchNorm statistics and the prior-dropout indices stay per rank
(as under DataParallel), the only exchange is a sum all-reduce of the flat fp32 gradient buffer, after which every rank
applies the same clip + Adam with ``grad_scale = 1/world`` — replicas stay bit-identical without parameter broadcasts.

These helpers are backend-agnostic (NCCL on the B200 box, gloo in the CPU tests).
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Initialise torch.distributed from torchrun's environment (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*).
    Returns (rank, local_rank, world).  A single process (no WORLD_SIZE) is world 1 and initialises nothing."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device('cuda', local))
        else:
            dist.init_process_group(backend)
    return rank, local, world


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of a global batch of n images for this rank (384 -> 8 x 48)."""
    if n % world != 0:
        raise ValueError(f'global batch {n} is not divisible by world size {world}')
    per = n // world
    return rank * per, (rank + 1) * per


def all_reduce_flat_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM all-reduce of a flat gradient buffer; the 1/world scaling is folded into the optimiser kernels."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def broadcast_flat_(flat: torch.Tensor, src: int = 0, group=None) -> torch.Tensor:
    """Step-0 parameter agreement (ranks construct from the same seed anyway; this makes it explicit)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(flat, src=src, group=group)
    return flat
