#!/usr/bin/env python
"""Timeline of CTA 0 of the tensor-core message-passing kernel (adj_apply_tc via xggm_adj_apply_fwd)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xggm_b200.functional as XF  # noqa: E402
from xggm_b200 import _lib  # noqa: E402
from xggm_b200._lib import call, ptr  # noqa: E402

B, N, H = 256, 36, 768
dev = torch.device("cuda")
x = torch.randn(B, N, H, device=dev)
adj = torch.rand(B, N, N, device=dev)
out = torch.empty_like(x)
wk = XF._adj_work(B, N, H, dev)
dbg = torch.zeros(32, dtype=torch.int64, device=dev)
lib = _lib.load()
for _ in range(3):
    call("xggm_adj_apply_fwd", ptr(adj), ptr(x), ptr(out), B, N, H, 1.0, None, 0.0, ptr(wk))
torch.cuda.synchronize()
lib.xggm_debug_timeline(dbg.data_ptr())
call("xggm_adj_apply_fwd", ptr(adj), ptr(x), ptr(out), B, N, H, 1.0, None, 0.0, ptr(wk))
torch.cuda.synchronize()
lib.xggm_debug_timeline(None)
t = dbg.cpu().tolist()
names = {0: "kernel start", 1: "setup done", 18: "kernel end"}
for i in range(4):
    names.update({2 + 4 * i: f"tile {i}: operands landed", 3 + 4 * i: f"tile {i}: last MMA issued",
                  4 + 4 * i: f"tile {i}: accumulator ready", 5 + 4 * i: f"tile {i}: epilogue done"})
for slot, ts in sorted(((s, v) for s, v in enumerate(t) if v), key=lambda kv: kv[1]):
    print(f"{(ts - t[0]) / 1e3:9.2f} us  {names.get(slot, slot)}")
