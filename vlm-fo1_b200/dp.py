"""Data-parallel plumbing: the path shards by independent images -- contiguous per-rank slices
of the sample list, full weight replica per GPU, and ONE collective per step: an all-gather of the decoded int32 ids
(NCCL over NVLink on GPUs; gloo in the CPU tests)."""
from __future__ import annotations

from typing import Tuple

import torch


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split: the first (n_total % world) ranks take one extra sample."""
    base, rem = divmod(n_total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_ids(tokens: torch.Tensor, lens: torch.Tensor, world: int, max_per_rank: int):
    """tokens int32 [b_local, T], lens int32 [b_local] -> (all_tokens [world*max_per_rank, T], all_lens [world*max_per_rank]).
    Ranks with fewer than max_per_rank samples pad with rows of -1 / length 0 (ragged batches)."""
    import torch.distributed as dist
    T = tokens.shape[1]
    pad_t = torch.full((max_per_rank, T), -1, dtype=torch.int32, device=tokens.device)
    pad_l = torch.zeros((max_per_rank,), dtype=torch.int32, device=tokens.device)
    pad_t[: tokens.shape[0]] = tokens
    pad_l[: lens.shape[0]] = lens
    if world == 1:
        return pad_t, pad_l
    out_t = torch.empty((world * max_per_rank, T), dtype=torch.int32, device=tokens.device)
    out_l = torch.empty((world * max_per_rank,), dtype=torch.int32, device=tokens.device)
    dist.all_gather_into_tensor(out_t, pad_t)
    dist.all_gather_into_tensor(out_l, pad_l)
    return out_t, out_l
