"""CUDA-graph capture of a whole GGM training step (forward + backward).

The block is ~100 short kernels per step; launched one by one from Python the step is bounded by
host launch overhead, not by the GPU.  ``GraphedStep`` records one call of a step function into a
``torch.cuda.CUDAGraph`` and replays it with a single launch:

    step = GraphedStep(fn, example_inputs)        # fn(*inputs) runs fwd + bwd, returns tensors
    outs = step(*new_inputs)                      # copies inputs into the static buffers, replays

Requirements on ``fn``: fixed shapes, no host synchronisation, gradients written into pre-allocated
buffers (``xggm_b200.ddp.FlatGrads``).  Randomness stays fresh across replays: torch's own generator
is graph-aware (``torch.randn``), and the library's dropout draws Philox bits keyed by a device-side
epoch counter that the captured step increments (``functional.dropout_epoch``).
"""
import itertools

import torch

from . import functional as XF

_graph_ids = itertools.count(1)


class GraphedStep:
    def __init__(self, fn, example_inputs, warmup=3):
        if not example_inputs or not all(t.is_cuda for t in example_inputs):
            raise RuntimeError("xggm_b200.GraphedStep: example inputs must be CUDA tensors")
        self.fn = fn
        self.static_in = [t.detach().clone() for t in example_inputs]
        dev = self.static_in[0].device
        self.epoch = torch.zeros(1, dtype=torch.int64, device=dev)
        self.site_base = next(_graph_ids) << 20   # keeps this graph's dropout sites apart from others'
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):             # warm-up off the capture: lazy init, allocator pools
            for _ in range(warmup):
                self._run()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = self._run()

    def _run(self):
        with XF.dropout_epoch(self.epoch, self.site_base):
            self.epoch.add_(1)
            return self.fn(*[t.detach() for t in self.static_in])

    # -- input prefetch: H2D of batch i+1 overlaps the replay of batch i ------------------------------
    def prefetch(self, *host_inputs):
        """Start copying the NEXT step's inputs (pinned host tensors) into device staging buffers on a
        side stream; returns immediately.  Pair with ``run_prefetched()``."""
        if len(host_inputs) != len(self.static_in):
            raise RuntimeError("xggm_b200.GraphedStep: wrong number of inputs")
        if not hasattr(self, "_stage"):
            self._stage = [torch.empty_like(t) for t in self.static_in]
            self._copy_stream = torch.cuda.Stream(device=self.static_in[0].device)
            self._staged = torch.cuda.Event()
            self._stage_free = torch.cuda.Event()
            self._stage_free.record(torch.cuda.current_stream(self.static_in[0].device))
        self._copy_stream.wait_event(self._stage_free)      # the previous batch has left the staging buffers
        with torch.cuda.stream(self._copy_stream):
            for s, t in zip(self._stage, host_inputs):
                s.copy_(t, non_blocking=True)
            self._staged.record(self._copy_stream)

    def run_prefetched(self):
        """Replay on the batch most recently passed to ``prefetch`` (device-to-device move into the
        static buffers, then one graph launch)."""
        cur = torch.cuda.current_stream(self.static_in[0].device)
        cur.wait_event(self._staged)
        for s, t in zip(self.static_in, self._stage):
            s.copy_(t, non_blocking=True)
        self._stage_free.record(cur)
        return self.replay()

    def replay(self):
        """Replay on the tensors already in ``static_in`` (no input copies)."""
        self.graph.replay()
        return self.static_out

    def __call__(self, *inputs):
        if len(inputs) != len(self.static_in):
            raise RuntimeError("xggm_b200.GraphedStep: wrong number of inputs")
        for s, t in zip(self.static_in, inputs):
            if t is not s:
                s.copy_(t, non_blocking=True)
        return self.replay()
