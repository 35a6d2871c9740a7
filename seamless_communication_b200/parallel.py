"""Data-parallel plumbing for the S2ST path: utterances are independent, so a global batch is split
contiguously across ranks (one process per GPU, full model replica each) and the only collectives are the scatter of
input waveforms from rank 0 and the gather of (padded) output waveforms / lengths back to it.  `torch.distributed`
(NCCL over NVLink on GPUs; gloo in the CPU tests) carries both - no collective ever follows a compute kernel on this
path, so there is nothing to fuse."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous static split of n items; the first n % world ranks get one extra item."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def scatter_batch(global_batch: Optional[torch.Tensor], per_rank: int, feat_shape, dtype, device, src: int = 0) -> torch.Tensor:
    """Rank `src` holds (world*per_rank, *feat_shape); every rank receives its (per_rank, *feat_shape) slice."""
    world, rank = dist.get_world_size(), dist.get_rank()
    recv = torch.empty((per_rank, *feat_shape), dtype=dtype, device=device)
    if rank == src:
        assert global_batch is not None and global_batch.shape[0] == world * per_rank
        dist.scatter(recv, [c.contiguous() for c in global_batch.to(device).chunk(world)], src=src)
    else:
        dist.scatter(recv, None, src=src)
    return recv


def gather_padded(local: torch.Tensor, lengths: torch.Tensor, dst: int = 0):
    """Gathers (B, T_r) tensors whose T_r may differ per rank: lengths first (to learn the max), then the data padded
    to the global max.  Returns (list of per-rank tensors trimmed to their own T_r, list of length tensors) on `dst`,
    (None, None) elsewhere."""
    world, rank = dist.get_world_size(), dist.get_rank()
    t_local = torch.tensor([local.shape[1]], dtype=torch.int64, device=local.device)
    t_all = [torch.zeros_like(t_local) for _ in range(world)]
    dist.all_gather(t_all, t_local)
    t_max = int(max(int(t.item()) for t in t_all))
    padded = torch.zeros((local.shape[0], t_max), dtype=local.dtype, device=local.device)
    padded[:, :local.shape[1]] = local
    outs = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    lens = [torch.empty_like(lengths) for _ in range(world)] if rank == dst else None
    dist.gather(padded, outs, dst=dst)
    dist.gather(lengths, lens, dst=dst)
    if rank != dst:
        return None, None
    return [o[:, :int(t.item())] for o, t in zip(outs, t_all)], lens
