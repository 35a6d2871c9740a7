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


class OverlappedExchange:
    """Host <-> device (<-> ranks) movement of a streaming S2ST job, double-buffered on side streams so that it runs under
    the neighbouring batches' compute (rank 0 owns the host buffers; NCCL only scatters inputs and gathers
    waveforms).  Per step:
        prefetch(host_global)  rank 0: pinned host -> HBM, then dist.scatter to every rank's `recv` buffer   [stream in]
        take()                 the compute stream waits for that batch and gets the waveform tensor
        publish(wavs, units)   pad the step's waveforms into one buffer; dist.gather to rank 0; HBM -> pinned   [stream out]
    With world == 1 the collectives drop out and only the copies remain.  Buffers are static (no allocator traffic
    across streams); `h2d_bytes` / `d2h_bytes` count what one step moves across PCIe on rank 0."""

    def __init__(self, per_rank: int, samples: int, max_out: int, device, world: int = 1, rank: int = 0, out_dtype=torch.float16):
        self.per_rank, self.samples, self.max_out, self.device, self.world, self.rank = per_rank, samples, max_out, device, world, rank
        self.s_in, self.s_out = torch.cuda.Stream(device), torch.cuda.Stream(device)
        self.recv = [torch.empty((per_rank, samples), dtype=torch.float32, device=device) for _ in range(2)]
        self.stage = ([torch.empty((world * per_rank, samples), dtype=torch.float32, device=device) for _ in range(2)]
                      if world > 1 and rank == 0 else None)
        self.ev_in = [torch.cuda.Event() for _ in range(2)]
        self.ev_out = [torch.cuda.Event() for _ in range(2)]
        self.ev_c = [torch.cuda.Event() for _ in range(2)]
        self.k_in = self.k_take = self.k_out = 0
        self.h2d_bytes = world * per_rank * samples * 4
        self.d2h_bytes = 0
        if max_out > 0:
            self.out = [torch.zeros((per_rank, max_out), dtype=out_dtype, device=device) for _ in range(2)]
            self.lens = [torch.zeros((per_rank,), dtype=torch.int32, device=device) for _ in range(2)]
            if rank == 0:
                self.gathered = ([torch.empty((world * per_rank, max_out), dtype=out_dtype, device=device) for _ in range(2)]
                                 if world > 1 else None)
                self.host_out = [torch.empty((world * per_rank, max_out), dtype=out_dtype).pin_memory() for _ in range(2)]
        self.results = []

    def prefetch(self, host_global: Optional[torch.Tensor]):
        k = self.k_in
        self.k_in ^= 1
        with torch.cuda.stream(self.s_in):
            if self.world == 1:
                self.recv[k].copy_(host_global, non_blocking=True)
            elif self.rank == 0:
                self.stage[k].copy_(host_global, non_blocking=True)
                dist.scatter(self.recv[k], list(self.stage[k].chunk(self.world)), src=0)
            else:
                dist.scatter(self.recv[k], None, src=0)
            self.ev_in[k].record(self.s_in)

    def take(self) -> torch.Tensor:
        k = self.k_take
        self.k_take ^= 1
        torch.cuda.current_stream().wait_event(self.ev_in[k])
        return self.recv[k]

    def publish(self, wavs: Optional[List[torch.Tensor]], units: Optional[List[List[int]]]):
        if self.max_out == 0 or wavs is None:
            return
        k = self.k_out
        self.k_out ^= 1
        main = torch.cuda.current_stream()
        main.wait_event(self.ev_out[k])  # the previous user of this buffer pair has left the device
        ns = []
        for i, w in enumerate(wavs):
            n = min(w.shape[-1], self.max_out)
            self.out[k][i, :n].copy_(w.reshape(-1)[:n])
            ns.append(n)
        self.lens[k].copy_(torch.tensor(ns, dtype=torch.int32), non_blocking=True)
        self.ev_c[k].record(main)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_c[k])
            if self.world > 1:
                dist.gather(self.out[k], list(self.gathered[k].chunk(self.world)) if self.rank == 0 else None, dst=0)
                src = self.gathered[k] if self.rank == 0 else None
            else:
                src = self.out[k]
            if self.rank == 0:
                self.host_out[k].copy_(src, non_blocking=True)
            self.ev_out[k].record(self.s_out)
        if self.rank == 0:
            self.d2h_bytes = self.host_out[k].numel() * self.host_out[k].element_size() + 8 * sum(len(u) for u in (units or []))

    def drain(self):
        self.s_in.synchronize()
        self.s_out.synchronize()
