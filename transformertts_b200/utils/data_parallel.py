"""Data-parallel plumbing of the training step: one process per GPU, torch.distributed (NCCL over NVLink on the GPUs,
gloo in the CPU tests).  The reference has no distributed code at all; this is new capability asked
for by BASELINE.json: the batch is sharded by rows, every rank runs the same forward/backward on its shard and the flat
fp32 gradient buffer is summed across ranks in a few contiguous buckets, each launched as soon as the backward pass has
finished writing it so the transfer overlaps the remaining backward kernels.  The 1/N factor is folded into Adam.
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None, device: Optional[torch.device] = None) -> Tuple[int, int]:
    """Initialise the default process group from torchrun's environment; returns (rank, world)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    if world > 1 and not dist.is_initialized():
        backend = backend or ('nccl' if torch.cuda.is_available() else 'gloo')
        kw = {}
        if backend == 'nccl' and device is not None:
            kw['device_id'] = device
        dist.init_process_group(backend, **kw)
    return rank, world


def shard_rows(n_rows: int, rank: int, world: int) -> slice:
    """Contiguous batch-row shard of rank `rank` (the first n_rows % world ranks get one extra row)."""
    base, extra = divmod(n_rows, world)
    start = rank * base + min(rank, extra)
    return slice(start, start + base + (1 if rank < extra else 0))


class GradSync:
    """Bucketed all-reduce of a flat gradient buffer.

    bucket_ready(lo, hi) may be called as soon as flat[lo:hi] is final (stream-ordered); finish() reduces whatever has not
    been sent yet, waits for everything and returns the scale (1/world) to apply to the summed gradient."""

    def __init__(self, flat: torch.Tensor, group=None):
        self.flat = flat
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._works: List = []
        self._sent: List[Tuple[int, int]] = []

    def bucket_ready(self, lo: int, hi: int):
        if self.world == 1 or hi <= lo:
            return
        self._works.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        self._sent.append((lo, hi))

    def finish(self) -> float:
        if self.world == 1:
            return 1.0
        pos = 0
        for lo, hi in sorted(self._sent):
            if lo > pos:
                self._works.append(dist.all_reduce(self.flat[pos:lo], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            pos = max(pos, hi)
        if pos < self.flat.numel():
            self._works.append(dist.all_reduce(self.flat[pos:], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        for w in self._works:
            w.wait()
        self._works, self._sent = [], []
        return 1.0 / self.world


class NativeGradSync:
    """The same contract as GradSync with the exchange done by libttsb's own NCCL entry points (include/ttsb.h: ttsb_dp_*):
    torch.distributed only carries the 128-byte NCCL id to the ranks once.  Buckets are reduced on a dedicated stream, ordered
    against the compute stream with two events, so a bucket overlaps the backward kernels that follow it."""

    _shared = {}   # device index -> (communicator, stream, world)

    def __init__(self, flat: torch.Tensor, group=None):
        from .. import lib
        self.lib = lib
        self.flat = flat
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._sent: List[Tuple[int, int]] = []
        if self.world == 1:
            return
        dev = flat.device.index
        if dev not in NativeGradSync._shared:
            rank = dist.get_rank(group)
            box = [lib.dp_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0, group=group)
            with torch.cuda.device(dev):
                comm = lib.dp_init(box[0], rank, self.world)
                NativeGradSync._shared[dev] = (comm, torch.cuda.Stream(device=dev), self.world)
        self.comm, self.stream, _ = NativeGradSync._shared[dev]

    def _reduce(self, lo: int, hi: int):
        cur = torch.cuda.current_stream(self.flat.device)
        ev = torch.cuda.Event()
        ev.record(cur)
        self.stream.wait_event(ev)          # the slice is final on the compute stream
        self.lib.dp_allreduce_bucket(self.comm, self.flat[lo:hi], self.stream.cuda_stream)

    def bucket_ready(self, lo: int, hi: int):
        if self.world == 1 or hi <= lo:
            return
        self._reduce(lo, hi)
        self._sent.append((lo, hi))

    def finish(self) -> float:
        if self.world == 1:
            return 1.0
        pos = 0
        for lo, hi in sorted(self._sent):
            if lo > pos:
                self._reduce(pos, lo)
            pos = max(pos, hi)
        if pos < self.flat.numel():
            self._reduce(pos, self.flat.numel())
        done = torch.cuda.Event()
        done.record(self.stream)
        torch.cuda.current_stream(self.flat.device).wait_event(done)   # Adam (compute stream) sees the reduced gradient
        self._sent = []
        return 1.0 / self.world


def make_grad_sync(flat: torch.Tensor, group=None):
    """GradSync implementation for this process: libttsb's NCCL path on CUDA (TTSB_DP_BACKEND=torch selects the
    torch.distributed one), torch.distributed (gloo) for CPU tensors in the host-logic tests."""
    backend = os.environ.get('TTSB_DP_BACKEND', 'native')
    if flat.is_cuda and backend == 'native' and dist.is_initialized() and dist.get_world_size(group) > 1:
        try:
            return NativeGradSync(flat, group)
        except Exception as e:   # NCCL not loadable: fall back, loudly
            print(f'[transformertts_b200] native NCCL path unavailable ({e}); using torch.distributed', flush=True)
    return GradSync(flat, group)


def global_loss(local_loss: torch.Tensor, local_numel: int, group=None) -> torch.Tensor:
    """The reference loss is a mean over the padded batch tensor; the single-process equivalent of a sharded batch is the
    numel-weighted mean of the shard losses."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_loss
    t = torch.stack([local_loss.detach().float() * local_numel, torch.tensor(float(local_numel), device=local_loss.device)])
    dist.all_reduce(t, group=group)
    return t[0] / t[1]
