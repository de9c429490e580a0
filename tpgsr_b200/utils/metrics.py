"""CTC greedy decode drop-in (reference: ``utils/metrics.py:71-88`` ``get_string_crnn``) on one CUDA kernel.

Same call surface as the reference — ``TextSR.eval`` binds it as the CRNN's ``string_process`` (``interfaces/super_resolution.py:115,134``) and
calls it on every recogniser output ``[T, B, C]`` (``:855,920,995,1022-1023``).  The reference copies the logits to the host and walks
``B x T`` elements in a Python loop of tensor indexing; here the device does arg-max, repeat collapse and blank removal
(``tpgsr_ctc_greedy_decode``) and the host fetches ``B x (T + 1)`` int32 in ONE copy and maps them through the alphabet."""
from __future__ import annotations

from typing import List, Tuple

import torch

from ..ops import _call, ptr, stream

ALPHABET = '-0123456789abcdefghijklmnopqrstuvwxyz'


def ctc_greedy_ids(outputs_: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Device part: logits ``[T, B, C]`` -> (ids ``[B, T]`` int32, decoded class indices then -1; lens ``[B]`` int32).  No host sync."""
    if not outputs_.is_cuda:
        raise RuntimeError('tpgsr_b200 metrics run on CUDA only — there is no CPU fallback')
    if outputs_.dim() != 3:
        raise ValueError(f'expected recogniser logits [T, B, C], got {tuple(outputs_.shape)}')
    x = outputs_.detach().float().contiguous()
    T, B, C = x.shape
    ids = torch.empty((B, T), dtype=torch.int32, device=x.device)
    lens = torch.zeros((B,), dtype=torch.int32, device=x.device)
    if B and T and C:
        _call('tpgsr_ctc_greedy_decode', ptr(x), T, B, C, ptr(ids), ptr(lens), stream())
    return ids, lens


def get_string_crnn(outputs_: torch.Tensor, alphabet: str = ALPHABET) -> List[str]:
    """``[T, B, C]`` recogniser logits -> list of B strings (class 0 = blank ``'-'``)."""
    if len(set(alphabet)) != len(alphabet):
        raise ValueError('get_string_crnn: the alphabet must consist of distinct characters (the reference collapses repeats by character)')
    T, B, C = outputs_.shape
    if C > len(alphabet):
        raise IndexError(f'get_string_crnn: {C} classes but the alphabet has {len(alphabet)} characters')
    if B == 0:
        return []
    ids, lens = ctc_greedy_ids(outputs_)
    host = torch.cat([lens.view(-1, 1), ids], 1).cpu().tolist()             # the only device->host transfer
    return [''.join(alphabet[i] for i in row[1:1 + row[0]]) for row in host]
