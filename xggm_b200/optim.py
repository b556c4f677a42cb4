"""Training-step tail of the reference trainers on fused kernels.

    optim = BertAdam(params, lr=..., warmup=0.1, t_total=...)       # src/lxrt/optimization.py:58-98
    loss = bce_with_logits(logit, target, scale=target.size(1))     # src/vqa/vqacpv2.py:110,173
    loss.backward()
    clip = clip_grad_norm_(optim, 5.)                                # src/vqa/vqacpv2.py:175
    optim.step(clip)                                                  # src/vqa/vqacpv2.py:176

The reference optimiser is a Python loop over parameters (three elementwise kernels per tensor, two more for
the clip); here every parameter group lives in four flat fp32 buffers (parameters, gradients -- the
``xggm_b200.ddp.FlatGrads`` bucket the backward kernels write and NCCL reduces --, first and second moments)
and a step is one gradient-norm reduction plus one update kernel per group (``xggm_grad_sumsq``,
``xggm_bertadam_step``).  ``clip_grad_norm_`` does not rescale the gradients in a pass of its own: it returns
the squared norm as a device scalar and the update kernel applies min(1, max_norm / (norm + 1e-6)) on the fly,
so the step needs no host synchronisation and can be captured in a CUDA graph.
"""
import ctypes as C
import math

import torch

from . import _lib
from ._lib import call, f32, ptr
from .ddp import FlatGrads


def warmup_cosine(x, warmup=0.002):
    return x / warmup if x < warmup else 0.5 * (1.0 + math.cos(math.pi * x))


def warmup_constant(x, warmup=0.002):
    return x / warmup if x < warmup else 1.0


def warmup_linear(x, warmup=0.002):
    return x / warmup if x < warmup else max((x - 1.0) / (warmup - 1.0), 0)


SCHEDULES = {"warmup_cosine": warmup_cosine, "warmup_constant": warmup_constant, "warmup_linear": warmup_linear}
_SCHED_ID = {"warmup_cosine": 0, "warmup_constant": 1, "warmup_linear": 2}   # XGGM_SCHED_* of include/xggm_b200.h


class _Group:
    def __init__(self, params, opts, flat_grads, lr_of=None):
        self.opts = opts
        self.lr_of = lr_of or {}      # id(parameter) -> base lr (several user groups sharing ONE bucket); else opts["lr"]
        self.grads = flat_grads if flat_grads is not None else FlatGrads(params)
        self.params = self.grads.params
        g = self.grads.flat
        # parameters move into one flat buffer with the offsets of the gradient bucket (in symmetric memory when the
        # gradient bucket is: the fused data-parallel step pushes parameter slices into the peers' buffers)
        self.param_ptrs = None
        flat_p = None
        if self.grads.peer_ptrs is not None:
            from .ddp import symmetric_empty
            flat_p, self.param_ptrs = symmetric_empty(g.numel(), g.dtype, g.device)
        self.flat_p = flat_p if flat_p is not None else torch.zeros_like(g)
        base = g.data_ptr()
        for p in self.params:
            o = (p.grad.data_ptr() - base) // g.element_size()
            view = self.flat_p[o:o + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
        self.m = torch.zeros_like(g)
        self.v = torch.zeros_like(g)
        # the optimiser is the only writer of these weights: their tensor-core operand planes may be cached and are
        # rebuilt by step() right after the update (functional.weight_planes)
        from . import functional as XF
        XF.cache_weight_planes([p for p in self.params if p.dim() == 2])
        # the step counter lives on the device (the update kernel evaluates the schedule from it and advances it,
        # so a CUDA-graph replay keeps following the schedule); `_ticket` is the kernel's hand-shake word
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=g.device)
        self._ticket = torch.zeros(1, dtype=torch.int32, device=g.device)

    def active_ranges(self):
        """(lo, hi, base_lr) element ranges of the bucket to update: parameters that received a gradient since the last
        zero_grad() (all of them if nothing was reported), contiguous ones merged while their base lr agrees."""
        fg = self.grads
        touched, out = fg._touched, []
        for i, (p, o) in enumerate(zip(fg.params, fg.offsets)):
            if touched and id(p) not in touched:
                continue
            hi = fg.offsets[i + 1] if i + 1 < len(fg.params) else fg.flat.numel()
            lr = self.lr_of.get(id(p), self.opts["lr"])
            if out and out[-1][1] == o and out[-1][2] == lr:
                out[-1] = (out[-1][0], hi, lr)
            else:
                out.append((o, hi, lr))
        return out

    @property
    def step(self):
        """Optimizer steps taken so far (reads the device counter: synchronises)."""
        return int(self.step_dev.item())

    def lr_scheduled(self, step=None):
        o = self.opts
        if o["t_total"] == -1:
            return o["lr"]
        step = self.step if step is None else step
        return o["lr"] * SCHEDULES[o["schedule"]](step / o["t_total"], o["warmup"])


class BertAdam:
    """Constructor signature and update rule of ``lxrt.optimization.BertAdam`` (no bias correction; decoupled
    weight decay; ``max_grad_norm`` accepted and, as in the reference, unused -- the trainers clip outside).

    ``params``: an iterable of parameters or of ``{"params": [...], "lr": ...}`` groups.  ``flat_grads``: an
    existing ``FlatGrads`` (or one per group) whose buffer already holds the parameters' gradients; otherwise
    one is created, which re-points every ``p.grad``.  Every ``p.data`` becomes a view of the group's flat
    parameter buffer.
    """

    def __init__(self, params, lr, warmup=-1, t_total=-1, schedule="warmup_linear", b1=0.9, b2=0.999, e=1e-6,
                 weight_decay=0.01, max_grad_norm=1.0, flat_grads=None):
        if lr < 0.0:
            raise ValueError("Invalid learning rate: {} - should be >= 0.0".format(lr))
        if schedule not in SCHEDULES:
            raise ValueError("Invalid schedule parameter: {}".format(schedule))
        if not 0.0 <= warmup < 1.0 and not warmup == -1:
            raise ValueError("Invalid warmup: {} - should be in [0.0, 1.0[ or -1".format(warmup))
        if not 0.0 <= b1 < 1.0:
            raise ValueError("Invalid b1 parameter: {} - should be in [0.0, 1.0[".format(b1))
        if not 0.0 <= b2 < 1.0:
            raise ValueError("Invalid b2 parameter: {} - should be in [0.0, 1.0[".format(b2))
        if not e >= 0.0:
            raise ValueError("Invalid epsilon value: {} - should be >= 0.0".format(e))
        defaults = dict(lr=lr, schedule=schedule, warmup=warmup, t_total=t_total, b1=b1, b2=b2, e=e,
                        weight_decay=weight_decay, max_grad_norm=max_grad_norm)
        params = list(params)
        if not params:
            raise ValueError("optimizer got an empty parameter list")
        groups = params if isinstance(params[0], dict) else [{"params": params}]
        if flat_grads is None or isinstance(flat_grads, FlatGrads):
            flat_grads = [flat_grads] * len(groups)
        if len(flat_grads) != len(groups):
            raise ValueError("need one FlatGrads per parameter group")
        self.groups = []
        self._user_groups = None
        shared = flat_grads[0] if (len(groups) > 1 and flat_grads[0] is not None and
                                   all(f is flat_grads[0] for f in flat_grads)) else None
        if shared is not None:
            # several parameter groups living in ONE bucket (the trainers' encoder / down-task groups, which differ in
            # lr only): one internal group with a per-parameter base lr, so that one fused data-parallel step -- one
            # joint gradient norm -- covers all of them (step_allreduce)
            lr_of, user = {}, []
            for grp in groups:
                extra = {k: v for k, v in grp.items() if k not in ("params", "lr")}
                if any(defaults.get(k) != v for k, v in extra.items()):
                    raise ValueError("parameter groups that share one FlatGrads may differ in lr only")
                plist = [p for p in grp["params"] if p.requires_grad]
                for p in plist:
                    lr_of[id(p)] = grp.get("lr", lr)
                user.append((dict(defaults, lr=grp.get("lr", lr)), plist))
            if {id(p) for p in shared.params} != set(lr_of):
                raise ValueError("flat_grads does not cover exactly the groups' parameters")
            self.groups.append(_Group(shared.params, dict(defaults), shared, lr_of))
            self._user_groups = user
            return
        for grp, fg in zip(groups, flat_grads):
            opts = dict(defaults)
            opts.update({k: v for k, v in grp.items() if k != "params"})
            plist = [p for p in grp["params"] if p.requires_grad]
            if fg is not None and {id(p) for p in fg.params} != {id(p) for p in plist}:
                raise ValueError("flat_grads does not cover exactly this group's parameters")
            self.groups.append(_Group(plist, opts, fg))

    @property
    def param_groups(self):
        if self._user_groups is not None:
            return [dict(o, params=pl) for o, pl in self._user_groups]
        return [dict(g.opts, params=g.params) for g in self.groups]

    def get_lr(self):
        """Scheduled learning rate of the NEXT step, one entry per parameter ([0] before the first step, as the
        reference does).  Reads the device step counters (synchronises)."""
        steps = [g.step for g in self.groups]
        if all(s == 0 for s in steps):
            return [0]
        out = []
        for g, s_ in zip(self.groups, steps):
            base = g.opts["lr"]
            for p in g.params:
                out.append(g.lr_scheduled(s_) * (g.lr_of.get(id(p), base) / base if base else 1.0))
        return out

    def zero_grad(self):
        """Zero the gradient buckets (keeps every ``p.grad`` linked; see ``FlatGrads.zero_``)."""
        for g in self.groups:
            g.grads.zero_()

    def step(self, clip=None):
        """One update of every group.  ``clip``: the handle returned by ``clip_grad_norm_`` (squared total
        gradient norm on the device + max_norm), applied inside the update kernel; None = no clipping.

        The learning-rate schedule is evaluated inside the kernel from a DEVICE step counter that the kernel
        advances, so the call is CUDA-graph capturable with any schedule (a host-computed lr would be baked into
        the capture).  As in the reference (src/lxrt/optimization.py:139-141, ``if p.grad is None: continue``)
        parameters that received no gradient since the last ``zero_grad()`` are left untouched -- no moment
        decay, no weight decay: the update runs over the bucket's active ranges only."""
        sumsq, max_norm = (None, 0.0) if clip is None else (clip.sumsq, clip.max_norm)
        for g in self.groups:
            o = g.opts
            _lib.check_device(g.flat_p)
            g.grads.relink()
            ranges = g.active_ranges()
            for i, (lo, hi, lr_r) in enumerate(ranges):
                sched = _lib.LrSchedule(g.step_dev.data_ptr(), g._ticket.data_ptr(), float(o["warmup"]),
                                        int(o["t_total"]), _SCHED_ID[o["schedule"]], int(i == len(ranges) - 1))
                call("xggm_bertadam_step_ex", ptr(g.flat_p[lo:hi]), ptr(g.grads.flat[lo:hi]), ptr(g.m[lo:hi]),
                     ptr(g.v[lo:hi]), hi - lo, float(lr_r), float(o["b1"]), float(o["b2"]), float(o["e"]),
                     float(o["weight_decay"]), ptr(sumsq), float(max_norm), C.cast(C.pointer(sched), C.c_void_p))
            if g.flat_p.is_cuda:
                from . import functional as XF
                XF.refresh_weight_planes(g.params)     # one launch: both operand layouts of every cached weight

    def fused_allreduce_available(self):
        """True when every group's buckets live in symmetric memory and there is ONE group (the squared norm the
        fused kernels clip by is the bucket's own)."""
        return len(self.groups) == 1 and self.groups[0].param_ptrs is not None and self.groups[0].grads.ctl_ptrs is not None

    def step_allreduce(self, max_norm=0.0, sumsq_out=None):
        """Data-parallel step: gradient all-reduce (average) + clip_grad_norm_(max_norm) + BertAdam update, fused
        over NVLink peer memory (xggm_dp_bertadam_step: reduce-scatter, 1/world of the update per rank, parameter
        all-gather by remote stores).  Replaces ``grads.all_reduce(); optim.step(clip_grad_norm_(grads, max_norm))``.
        Every rank must call it once per step.  Returns the device scalar holding the squared total norm."""
        if not self.fused_allreduce_available():
            raise RuntimeError("xggm_b200.BertAdam.step_allreduce needs one parameter group whose FlatGrads was built "
                               "with symmetric=True under an NCCL process group (world > 1)")
        import torch.distributed as dist
        g = self.groups[0]
        o = g.opts
        _lib.check_device(g.flat_p)
        g.grads.relink()
        ranges = g.active_ranges()
        if len(ranges) > _lib.DP_MAX_RANGES:
            raise RuntimeError(f"xggm_b200.BertAdam.step_allreduce: {len(ranges)} active ranges (max {_lib.DP_MAX_RANGES}); "
                               "lay parameters with a common lr out contiguously")
        peers = _lib.DpPeers()
        peers.rank, peers.world = dist.get_rank(), dist.get_world_size()
        for k in range(peers.world):
            peers.grad[k], peers.param[k], peers.ctl[k] = g.grads.peer_ptrs[k], g.param_ptrs[k], g.grads.ctl_ptrs[k]
        gmc, pmc = getattr(g.grads.flat, "_xggm_multicast_ptr", 0), getattr(g.flat_p, "_xggm_multicast_ptr", 0)
        peers.grad_multicast = gmc if (gmc and pmc) else None      # in-switch reduce / broadcast when the fabric offers it
        peers.param_multicast = pmc if (gmc and pmc) else None
        lo = (C.c_longlong * len(ranges))(*[r[0] for r in ranges])
        hi = (C.c_longlong * len(ranges))(*[r[1] for r in ranges])
        lrs = (C.c_double * len(ranges))(*[float(r[2]) for r in ranges])
        sched = _lib.LrSchedule(g.step_dev.data_ptr(), g._ticket.data_ptr(), float(o["warmup"]), int(o["t_total"]),
                                _SCHED_ID[o["schedule"]], 1)
        if sumsq_out is None:
            sumsq_out = torch.empty(1, device=g.flat_p.device, dtype=torch.float32)
        call("xggm_dp_bertadam_step", C.cast(C.pointer(peers), C.c_void_p), ptr(g.m), ptr(g.v), g.flat_p.numel(), lo, hi, lrs,
             len(ranges), float(o["lr"]), float(o["b1"]), float(o["b2"]), float(o["e"]), float(o["weight_decay"]),
             float(max_norm), C.cast(C.pointer(sched), C.c_void_p), ptr(sumsq_out))
        from . import functional as XF
        XF.refresh_weight_planes(g.params)
        return GradClip(sumsq_out, max_norm)

    # -- checkpointing (the reference trainers do not save optimiser state; torch-style layout) ----
    def state_dict(self):
        """``{"state": {index: {"step", "next_m", "next_v"}}, "param_groups": [...]}`` with the reference's state
        keys (src/lxrt/optimization.py:146-152); parameter indices run over the groups in order."""
        state, groups, idx = {}, [], 0
        for g in self.groups:
            step = g.step
            ids = []
            for p, o in zip(g.params, g.grads.offsets):
                n = p.numel()
                state[idx] = {"step": step, "next_m": g.m[o:o + n].view_as(p).clone(),
                              "next_v": g.v[o:o + n].view_as(p).clone()}
                ids.append(idx)
                idx += 1
            groups.append(dict(g.opts, params=ids))
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd):
        idx = 0
        if len(sd["param_groups"]) != len(self.groups):
            raise ValueError("loaded state dict has a different number of parameter groups")
        for g, saved in zip(self.groups, sd["param_groups"]):
            if len(saved["params"]) != len(g.params):
                raise ValueError("loaded state dict contains a parameter group that doesn't match the size of "
                                 "optimizer's group")
            g.opts.update({k: v for k, v in saved.items() if k != "params"})
            step = 0
            for p, o in zip(g.params, g.grads.offsets):
                st = sd["state"].get(idx, sd["state"].get(str(idx)))
                if st is not None:
                    n = p.numel()
                    g.m[o:o + n].copy_(st["next_m"].reshape(-1))
                    g.v[o:o + n].copy_(st["next_v"].reshape(-1))
                    step = max(step, int(st["step"]))
                idx += 1
            g.step_dev.fill_(step)


class GradClip:
    """Squared total gradient norm (device scalar, no host sync) and the clip threshold."""

    def __init__(self, sumsq, max_norm):
        self.sumsq, self.max_norm = sumsq, float(max_norm)

    def total_norm(self):
        """The value torch.nn.utils.clip_grad_norm_ returns (synchronises)."""
        return float(self.sumsq.sqrt())


def clip_grad_norm_(source, max_norm, sumsq=None):
    """torch.nn.utils.clip_grad_norm_(parameters, max_norm) for gradients held in flat buckets.

    ``source``: a ``BertAdam``, a ``FlatGrads`` or a list of them -- everything that is clipped TOGETHER (the
    reference clips ``self.model.parameters()``, i.e. the block and the encoder, by their joint norm; pass
    ``sumsq`` to keep accumulating into an existing scalar).  Returns a ``GradClip`` for ``BertAdam.step``.
    """
    if isinstance(source, BertAdam):
        buckets = [g.grads for g in source.groups]
    elif isinstance(source, FlatGrads):
        buckets = [source]
    else:
        buckets = [b for s in source for b in ([g.grads for g in s.groups] if isinstance(s, BertAdam) else [s])]
    if not buckets:
        raise ValueError("nothing to clip")
    accumulate = sumsq is not None
    if sumsq is None:
        sumsq = torch.empty(1, device=buckets[0].flat.device, dtype=torch.float32)
    for b in buckets:
        flat = f32(b.flat, "gradient bucket")
        call("xggm_grad_sumsq", ptr(flat), flat.numel(), ptr(sumsq), int(accumulate))
        accumulate = True
    return GradClip(sumsq, max_norm)


class _BceLogits(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logit, target, scale):
        logit, target = f32(logit, "logit"), f32(target, "target")
        if logit.shape != target.shape:
            raise RuntimeError("xggm_b200.bce_with_logits: shape mismatch")
        loss = torch.empty(1, device=logit.device, dtype=torch.float32)
        call("xggm_bce_logits_fwd", ptr(logit), ptr(target), float(scale), ptr(loss), logit.numel())
        ctx.save_for_backward(logit, target)
        ctx.scale = float(scale)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        logit, target = ctx.saved_tensors
        gl = torch.empty_like(logit)
        call("xggm_bce_logits_bwd", ptr(logit), ptr(target), ptr(f32(g).reshape(1)), ctx.scale, ptr(gl), logit.numel())
        return gl, None, None


def bce_with_logits(logit, target, scale=1.0):
    """nn.BCEWithLogitsLoss()(logit, target) * scale (the trainers use scale = target.size(1))."""
    return _BceLogits.apply(logit, target, scale)
