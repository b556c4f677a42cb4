"""Data-parallel plumbing for the graph block (one process per GPU, NCCL; gloo in CPU tests).

Every graph is independent, so ranks take disjoint shards of the batch and
the only exchange is the gradient all-reduce.  All parameter gradients live in ONE flat fp32
buffer that the backward kernels write in place, so a step needs a single collective and no
staging copy.  The reference's branch choice (``random.randint(1, 10) <= delta``,
src/vqa/vqacpv2.py:192-194) must be identical on every rank or the ranks would reduce different
parameter sets: ``BranchSchedule`` derives it from a shared seed.
"""
import os
import random

import torch
import torch.distributed as dist


def symmetric_empty(numel, dtype, device, zero=True):
    """A tensor every rank of the default process group can address through NVLink / NVSwitch peer pointers
    (torch symmetric memory: CUDA VMM allocations exchanged at rendezvous).  Returns (tensor, peer_pointers) -- the
    base address of the same allocation on every rank, this process's mapping -- or (None, None) when symmetric
    memory is not available (no NCCL group, one rank, or an older torch)."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1) or dist.get_backend() != "nccl":
        return None, None
    try:
        import torch.distributed._symmetric_memory as symm_mem
        group = dist.group.WORLD
        try:
            symm_mem.enable_symm_mem_for_group(group.group_name)
        except Exception:
            pass
        t = symm_mem.empty(numel, dtype=dtype, device=device)
        if zero:
            # zero BEFORE the rendezvous: the rendezvous is a collective, so once any rank leaves it every rank's buffer is
            # already clean -- zeroing afterwards could wipe a flag a faster peer has just written into it
            t.zero_()
            torch.cuda.synchronize(device)
        hdl = symm_mem.rendezvous(t, group.group_name)
        dist.barrier()
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        t._xggm_symm_handle = hdl          # keeps the mapping alive
        mc = 0
        try:                               # NVSwitch multicast (NVLS) address of the same allocation, if the fabric has one
            # measured (B=256, 33 MB bucket): 8 GPUs 1.962 ms with multicast vs 1.981 unicast; 2 GPUs 1.971 vs 1.927 -- the
            # in-switch reduction pays off once a rank would otherwise read more than a couple of peers
            want = os.environ.get("XGGM_DP_MULTICAST", "auto")
            use = want == "1" or (want == "auto" and dist.get_world_size() > 2)
            if use and getattr(hdl, "has_multicast_support", False):
                mc = int(hdl.multicast_ptr or 0)
        except Exception:
            mc = 0
        t._xggm_multicast_ptr = mc
        return t, ptrs
    except Exception as e:                 # pragma: no cover - depends on the platform
        import warnings
        warnings.warn(f"xggm_b200: symmetric memory unavailable ({e!r}); using NCCL collectives")
        return None, None


def shard_range(n_items, rank, world):
    """Contiguous shard [lo, hi) of ``n_items`` for ``rank``; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class FlatGrads:
    """Point every ``p.grad`` at a slice of one flat buffer (same dtype/device as the parameters).

    ``early``: parameters whose gradients are complete first in the backward pass (the modules applied LAST
    in the forward pass: fusion_fc and the last generator layer).  They are laid out at the front of the
    buffer, so ``reduce_early()`` -- fired by ``overlap()`` when backward crosses the boundary between the
    last two generator layers -- can put that bucket on the wire on a side stream while the rest of the
    backward pass is still computing; ``all_reduce()`` then only has the remaining bucket left.
    """

    def __init__(self, params, early=(), symmetric=False):
        """symmetric=True (NCCL group, world > 1): the bucket lives in NVLink-addressable symmetric memory so that
        ``BertAdam.step_allreduce`` can run the fused reduce-scatter + clip + update + all-gather kernels on it."""
        params = [p for p in params if p.requires_grad]
        early_ids = {id(p) for p in early}
        self.params = [p for p in params if id(p) in early_ids] + [p for p in params if id(p) not in early_ids]
        if not self.params:
            raise ValueError("no trainable parameters")
        p0 = self.params[0]
        # every slice starts on a 128-byte boundary: the kernels' vector (float4) paths need 16-byte alignment
        align = 128 // p0.element_size()
        offs, off = [], 0
        self.split = 0
        for p in self.params:
            offs.append(off)
            off += (p.numel() + align - 1) // align * align
            if id(p) in early_ids:
                self.split = off
        self.peer_ptrs = self.ctl = self.ctl_ptrs = None
        flat = None
        if symmetric and p0.is_cuda and p0.dtype == torch.float32:
            flat, self.peer_ptrs = symmetric_empty(off, p0.dtype, p0.device)
            if flat is not None:
                self.ctl, self.ctl_ptrs = symmetric_empty(128, torch.int32, p0.device)   # XGGM_DP_CTL_BYTES
                if self.ctl is None:
                    flat, self.peer_ptrs = None, None
        self.flat = flat if flat is not None else torch.zeros(off, device=p0.device, dtype=p0.dtype)
        self.offsets = offs
        for p, o in zip(self.params, offs):
            p.grad = self.flat[o:o + p.numel()].view_as(p)
            # opt-in marker for functional.FUSE_GRAD_ACCUMULATION: only parameters registered here may have
            # their gradients accumulated in place by the backward kernels
            p._xggm_flat = self
            if hasattr(p, "register_post_accumulate_grad_hook"):   # gradients that arrive through autograd itself
                p.register_post_accumulate_grad_hook(lambda q, self=self: self.touch(q))
        self._touched = set()
        self._side = None
        self._early_in_flight = False
        self._average = True

    def zero_(self):
        """Zero the bucket (use this, or ``BertAdam.zero_grad()``, instead of ``model.zero_grad()``: the latter
        sets every ``p.grad`` to None on torch >= 2 and un-links the parameters from the bucket; ``relink()``
        -- called by ``all_reduce`` and ``BertAdam.step`` -- repairs that, at the price of a copy)."""
        self.flat.zero_()
        self._touched.clear()

    # -- which parameters received a gradient since the last zero_() ------------------------------------
    def touch(self, p):
        self._touched.add(id(p))

    def active_ranges(self):
        """Contiguous [lo, hi) element ranges of the bucket that belong to parameters whose gradient was
        written since the last ``zero_()`` (the reference optimiser skips parameters whose ``.grad`` is None,
        src/lxrt/optimization.py:139-141: e.g. ``encoder_adj`` in a node-branch step).  If nothing was reported
        (gradients written by other means) every parameter counts as active."""
        if not self._touched:
            return [(0, self.flat.numel())]
        ranges = []
        for i, (p, o) in enumerate(zip(self.params, self.offsets)):
            if id(p) not in self._touched:
                continue
            hi = self.offsets[i + 1] if i + 1 < len(self.params) else self.flat.numel()
            if ranges and ranges[-1][1] == o:
                ranges[-1] = (ranges[-1][0], hi)
            else:
                ranges.append((o, hi))
        return ranges

    def relink(self, strict=False):
        """Make sure every ``p.grad`` is still its slice of the bucket.  ``model.zero_grad()`` (set_to_none) or
        code that assigns ``p.grad`` breaks the link: the next backward then allocates gradients OUTSIDE the
        bucket and the all-reduce / optimiser would silently read stale zeros.  Stray gradients are copied into
        their slice and the link is restored (``strict=True`` raises instead).  Returns the number of repairs."""
        base, esz, fixed = self.flat.data_ptr(), self.flat.element_size(), 0
        for p, o in zip(self.params, self.offsets):
            g = p.grad
            want = base + o * esz
            if g is not None and g.data_ptr() == want and g.shape == p.shape and g.is_contiguous():
                continue
            if strict:
                raise RuntimeError("xggm_b200.FlatGrads: a parameter's .grad no longer points into the flat bucket "
                                   "(model.zero_grad() sets it to None on torch >= 2); use FlatGrads.zero_() / "
                                   "BertAdam.zero_grad(), or call relink()")
            view = self.flat[o:o + p.numel()].view_as(p)
            if g is not None:
                view.copy_(g)
                self.touch(p)
            p.grad = view
            fixed += 1
        return fixed

    @staticmethod
    def _distributed():
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def _reduce(self, buf, average):
        if average and dist.get_backend() == "nccl":
            dist.all_reduce(buf, op=dist.ReduceOp.AVG)     # averaged inside the collective: no extra pass
            return
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        if average:
            buf.div_(dist.get_world_size())

    def reduce_early(self):
        """All-reduce the early bucket now, on a side stream that waits for the work enqueued so far."""
        if not self._distributed() or self.split == 0 or self._early_in_flight:
            return
        if self.flat.is_cuda:
            cur = torch.cuda.current_stream(self.flat.device)
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.flat.device)
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                self._reduce(self.flat[:self.split], self._average)
        else:
            self._reduce(self.flat[:self.split], self._average)
        self._early_in_flight = True

    def all_reduce(self, average=True):
        """Sum (or average) the gradients over all ranks: one collective, or -- when ``reduce_early()`` already
        shipped the first bucket during the backward pass -- the remaining bucket plus a join."""
        self.relink()
        if not self._distributed():
            return self.flat
        if self._early_in_flight:
            if average != self._average:
                raise RuntimeError("xggm_b200.FlatGrads: overlap(average=...) and all_reduce(average=...) disagree")
            self._reduce(self.flat[self.split:], average)
            if self.flat.is_cuda:
                torch.cuda.current_stream(self.flat.device).wait_stream(self._side)
            self._early_in_flight = False
            return self.flat
        self._reduce(self.flat, average)
        return self.flat

    def overlap(self, average=True):
        """Context manager for one forward+backward: generators report the boundary between their last two
        layers (``notify_layer_boundary``); when backward reaches it the early bucket is reduced."""
        return _Overlap(self, average)


_active = None   # the FlatGrads whose overlap() context is open (one per process: one process per GPU)


class _Overlap:
    def __init__(self, grads, average):
        self.grads, self.average = grads, average

    def __enter__(self):
        global _active
        self.grads._average = self.average
        _active = self.grads
        return self.grads

    def __exit__(self, *exc):
        global _active
        _active = None


def notify_layer_boundary(x):
    """Called by the generators on the tensor that enters their LAST layer.  Its gradient exists exactly when
    the backward pass of everything after it (last layer, read-out, losses) has finished."""
    g = _active
    if g is not None and g.split > 0 and g._distributed() and isinstance(x, torch.Tensor) and x.requires_grad:
        def _fire(_grad, g=g):
            g.reduce_early()
        x.register_hook(_fire)
    return x


class BranchSchedule:
    """Rank-synchronous replacement of the trainers' ``random.randint(1, 10)`` branch draw."""

    def __init__(self, delta, seed=9595):  # 9595 = the reference's default seed, src/param.py:49
        self.delta = delta
        self.rng = random.Random(seed)

    def next(self):
        """'relation' if r <= delta else 'node' (src/vqa/vqacpv2.py:193-194,226)."""
        return "relation" if self.rng.randint(1, 10) <= self.delta else "node"
