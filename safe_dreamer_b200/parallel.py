"""Data-parallel plumbing for the RSSM hot path: one process per GPU, the replay batch
is sharded over ranks, every rank runs the identical single-GPU scans on its slice, and the only exchange
step is the all-reduce (mean) of the RSSM weight gradients that `sd_observe_bwd` accumulates.

The gradients live in ONE flat fp32 buffer (the per-parameter tensors handed to the C ABI are views into
it), so the exchange is a single NCCL all-reduce with no packing copies; it is issued asynchronously right
after the backward scan and overlaps the imagination rollout, which does not depend on it.
The reference has no distributed code at all; semantics follow standard DDP: per-replica
batch statistics, gradients averaged over ranks."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_rows(global_rows: int, rank: int, world: int):
    """Contiguous row range [lo, hi) of `rank`; remainders go to the lowest ranks."""
    base, rem = divmod(global_rows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class GradBucket:
    """Flat gradient buffer with named views + asynchronous mean all-reduce."""

    def __init__(self, shapes: dict, device, dtype=torch.float32):
        self.names = list(shapes)
        sizes = [int(torch.Size(shapes[n]).numel()) for n in self.names]
        self.flat = torch.zeros(sum(sizes), dtype=dtype, device=device)
        self.views, off = {}, 0
        for n, sz in zip(self.names, sizes):
            self.views[n] = self.flat[off:off + sz].view(shapes[n])
            off += sz
        self._work = None

    def zero_(self):
        self.flat.zero_()

    def allreduce_async(self, group=None):
        """Launch the mean all-reduce; returns immediately (NCCL runs on its own stream, ordered after the
        work already enqueued on the current stream)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            self._work = None
            return
        # NCCL averages inside the collective (no extra pass over the 24 MB bucket afterwards); gloo (the CPU tests) has no
        # AVG, so there the 1/world scale is applied in wait()
        avg = dist.get_backend(group) == "nccl"
        self._scale = None if avg else 1.0 / dist.get_world_size(group)
        self._work = dist.all_reduce(self.flat, op=dist.ReduceOp.AVG if avg else dist.ReduceOp.SUM, group=group, async_op=True)

    def wait(self):
        """Make the current stream wait for the exchange (and apply the 1/world scale where the backend did not)."""
        if self._work is not None:
            self._work.wait()
            if self._scale is not None:
                self.flat.mul_(self._scale)
            self._work = None
