"""Multi-GPU plumbing for the hot path: utterances are independent, so they shard across
ranks with NO data-path collective; the one collective is the gather that re-joins decoded waveforms.
One process per GPU, torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


def shard_utterances(n_utterances: int, rank: int, world: int) -> List[int]:
    """Round-robin assignment: rank r takes utterances r, r + world, ... (SURVEY.md 8e "Partitioning")."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_utterances, world))


def gather_waveforms(local: torch.Tensor, lengths: torch.Tensor, n_utterances: int, group=None, stream=None):
    """local [B_local, T_max] waveforms of the utterances `shard_utterances` gave this rank (rows padded to the same
    T_max on every rank), lengths [B_local] int64.  Returns (waves [n_utterances, T_max], lens [n_utterances]) in
    ORIGINAL utterance order on every rank.  Ranks with fewer utterances pad with empty rows so a single
    all_gather_into_tensor (the only collective of the path) suffices."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per_rank = (n_utterances + world - 1) // world
    b_local, t_max = local.shape
    if b_local != len(shard_utterances(n_utterances, rank, world)):
        raise ValueError("local batch does not match this rank's shard")
    buf = torch.zeros((per_rank, t_max), dtype=local.dtype, device=local.device)
    lbuf = torch.zeros(per_rank, dtype=torch.int64, device=local.device)
    buf[:b_local] = local
    lbuf[:b_local] = lengths.to(local.device)
    out = torch.empty((world * per_rank, t_max), dtype=local.dtype, device=local.device)
    lout = torch.empty(world * per_rank, dtype=torch.int64, device=local.device)
    ctx = torch.cuda.stream(stream) if stream is not None else _null()
    with ctx:
        dist.all_gather_into_tensor(out, buf, group=group)
        dist.all_gather_into_tensor(lout, lbuf, group=group)
    # rank r's row j is utterance r + j*world
    order = [r * per_rank + j for u in range(n_utterances) for r, j in [(u % world, u // world)]]
    idx = torch.as_tensor(order, device=local.device)
    return out[idx], lout[idx]


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
