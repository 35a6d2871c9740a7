"""Data-parallel plumbing for the S2ST path: utterances are independent, so a global batch is split
contiguously across ranks (one process per GPU, full model replica each) and the only collectives are the scatter of
input waveforms from rank 0 and the gather of (padded) output waveforms / lengths back to it.  `torch.distributed`
(NCCL over NVLink on GPUs; gloo in the CPU tests) carries both - no collective ever follows a compute kernel on this
path, so there is nothing to fuse."""
from __future__ import annotations

import concurrent.futures
import queue
import threading
from typing import Callable, List, Optional, Sequence, Tuple

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

    def __init__(self, per_rank: int, samples: int, max_out: int, device, world: int = 1, rank: int = 0, out_dtype=torch.float16,
                 slots: int = 2):
        """`slots`: depth of the buffer rings (2 = double buffering; LanePool users need lanes + 2 so that a batch still
        being read by a lane is never overwritten by a prefetch)."""
        self.per_rank, self.samples, self.max_out, self.device, self.world, self.rank = per_rank, samples, max_out, device, world, rank
        self.slots = slots
        self.s_in, self.s_out = torch.cuda.Stream(device), torch.cuda.Stream(device)
        self.recv = [torch.empty((per_rank, samples), dtype=torch.float32, device=device) for _ in range(slots)]
        self.stage = ([torch.empty((world * per_rank, samples), dtype=torch.float32, device=device) for _ in range(slots)]
                      if world > 1 and rank == 0 else None)
        self.ev_in = [torch.cuda.Event() for _ in range(slots)]
        self.ev_out = [torch.cuda.Event() for _ in range(slots)]
        self.ev_c = [torch.cuda.Event() for _ in range(slots)]
        self.ev_free: List[Optional[torch.cuda.Event]] = [None] * slots  # set by release(): the consumer of recv[k] is done
        self.k_in = self.k_take = self.k_out = 0
        self.h2d_bytes = world * per_rank * samples * 4
        self.d2h_bytes = 0
        if max_out > 0:
            self.out = [torch.zeros((per_rank, max_out), dtype=out_dtype, device=device) for _ in range(slots)]
            self.lens = [torch.zeros((per_rank,), dtype=torch.int32, device=device) for _ in range(slots)]
            if rank == 0:
                self.gathered = ([torch.empty((world * per_rank, max_out), dtype=out_dtype, device=device) for _ in range(slots)]
                                 if world > 1 else None)
                self.host_out = [torch.empty((world * per_rank, max_out), dtype=out_dtype).pin_memory() for _ in range(slots)]
        self.results = []

    def prefetch(self, host_global: Optional[torch.Tensor]):
        k = self.k_in
        self.k_in = (k + 1) % self.slots
        with torch.cuda.stream(self.s_in):
            if self.ev_free[k] is not None:
                self.s_in.wait_event(self.ev_free[k])
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
        self.k_take = (k + 1) % self.slots
        torch.cuda.current_stream().wait_event(self.ev_in[k])
        return self.recv[k]

    def take_async(self) -> Tuple[torch.Tensor, torch.cuda.Event, int]:
        """For a consumer on another stream (LanePool): (batch, the event to wait for, slot to `release` afterwards)."""
        k = self.k_take
        self.k_take = (k + 1) % self.slots
        return self.recv[k], self.ev_in[k], k

    def release(self, k: int, done: torch.cuda.Event):
        self.ev_free[k] = done

    def publish(self, wavs: Optional[List[torch.Tensor]], units: Optional[List[List[int]]]):
        if self.max_out == 0 or wavs is None:
            return
        k = self.k_out
        self.k_out = (k + 1) % self.slots
        main = torch.cuda.current_stream()
        main.wait_event(self.ev_out[k])  # the previous user of this buffer pair has left the device
        ns = []
        for i, w in enumerate(wavs):
            w.record_stream(main)  # produced on a lane's stream, read here
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


class LanePool:
    """`lanes` independent batches in flight on one GPU.

    The S2ST beam search is a chain of ~260 short, latency-bound kernels per step that leaves most of the SMs and of the
    HBM bandwidth idle (bench.py `roofline`), while batches are independent of one another.  Each lane is a
    host thread bound to its own CUDA stream and to its own search-state lane of every engine in `engines`
    (engine.set_lane: KV caches, history and captured step graphs are per lane), so the hardware interleaves the lanes'
    kernel chains; the encoder / T2U / vocoder GEMMs of one batch fill the gaps of another batch's search.
    `submit(lane, fn, *args, after=event)` runs fn on that lane's thread and stream (after `event`, if given) and returns
    a Future of (result, completion event).  Jobs of one lane run in submission order."""

    def __init__(self, device, lanes: int, engines: Sequence = ()):
        self.device, self.lanes, self.engines = torch.device(device), int(lanes), list(engines)
        self.cuda = self.device.type == "cuda"  # on "cpu" (host-logic tests) the lanes are plain threads: no streams / events
        if self.cuda and self.device.index is None:  # "cuda" -> the caller's current device (the workers need an explicit index)
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.streams = [torch.cuda.Stream(self.device) if self.cuda else None for _ in range(self.lanes)]
        self.queues = [queue.SimpleQueue() for _ in range(self.lanes)]
        self.threads = [threading.Thread(target=self._run, args=(i,), daemon=True, name=f"lane{i}") for i in range(self.lanes)]
        for t in self.threads:
            t.start()

    def _run(self, i: int):
        setup_error = None
        try:
            if self.cuda:
                torch.cuda.set_device(self.device)
            for e in self.engines:
                e.set_lane(i)
        except BaseException as e:  # every job of this lane then fails with it instead of waiting forever
            setup_error = e
        s = self.streams[i]
        while True:
            item = self.queues[i].get()
            if item is None:
                return
            fn, args, after, fut = item
            if setup_error is not None:
                fut.set_exception(setup_error)
                continue
            try:
                if not self.cuda:
                    with torch.inference_mode():
                        fut.set_result((fn(*args), None))
                    continue
                with torch.inference_mode(), torch.cuda.stream(s):
                    if after is not None:
                        s.wait_event(after)
                    out = fn(*args)
                    ev = torch.cuda.Event()
                    ev.record(s)
                fut.set_result((out, ev))
            except BaseException as e:  # delivered to the caller through the future
                fut.set_exception(e)

    def submit(self, lane: int, fn: Callable, *args, after: Optional[torch.cuda.Event] = None) -> concurrent.futures.Future:
        fut: concurrent.futures.Future = concurrent.futures.Future()
        self.queues[lane % self.lanes].put((fn, args, after, fut))
        return fut

    def warm(self, fn: Callable, *args):
        """Run fn once on every lane, one lane at a time (first use captures the lane's CUDA graphs)."""
        for i in range(self.lanes):
            self.submit(i, fn, *args).result()
        if self.cuda:
            torch.cuda.synchronize(self.device)

    def map(self, fn: Callable, arg_lists: Sequence[Sequence]) -> list:
        """fn(*args) for every args in arg_lists, round-robin over the lanes; results in order, device work complete."""
        futs = [self.submit(i, fn, *a) for i, a in enumerate(arg_lists)]
        outs = []
        for f in futs:
            out, ev = f.result()
            if ev is not None:
                ev.synchronize()
            outs.append(out)
        return outs

    def close(self):
        for q in self.queues:
            q.put(None)
        for t in self.threads:
            t.join(timeout=10)
