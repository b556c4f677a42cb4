"""Data-parallel plumbing for the graph block (one process per GPU, NCCL; gloo in CPU tests).

Every graph is independent, so ranks take disjoint shards of the batch and
the only exchange is the gradient all-reduce.  All parameter gradients live in ONE flat fp32
buffer that the backward kernels write in place, so a step needs a single collective and no
staging copy.  The reference's branch choice (``random.randint(1, 10) <= delta``,
src/vqa/vqacpv2.py:192-194) must be identical on every rank or the ranks would reduce different
parameter sets: ``BranchSchedule`` derives it from a shared seed.
"""
import random

import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous shard [lo, hi) of ``n_items`` for ``rank``; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class FlatGrads:
    """Point every ``p.grad`` at a slice of one flat buffer (same dtype/device as the parameters)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        p0 = self.params[0]
        # every slice starts on a 128-byte boundary: the kernels' vector (float4) paths need 16-byte alignment
        align = 128 // p0.element_size()
        offs, off = [], 0
        for p in self.params:
            offs.append(off)
            off += (p.numel() + align - 1) // align * align
        self.flat = torch.zeros(off, device=p0.device, dtype=p0.dtype)
        for p, o in zip(self.params, offs):
            p.grad = self.flat[o:o + p.numel()].view_as(p)

    def zero_(self):
        self.flat.zero_()

    def all_reduce(self, average=True):
        """Sum (or average) the gradients over all ranks with one collective."""
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
            return self.flat
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        if average:
            self.flat.div_(dist.get_world_size())
        return self.flat


class BranchSchedule:
    """Rank-synchronous replacement of the trainers' ``random.randint(1, 10)`` branch draw."""

    def __init__(self, delta, seed=9595):  # 9595 = the reference's default seed, src/param.py:49
        self.delta = delta
        self.rng = random.Random(seed)

    def next(self):
        """'relation' if r <= delta else 'node' (src/vqa/vqacpv2.py:193-194,226)."""
        return "relation" if self.rng.randint(1, 10) <= self.delta else "node"
