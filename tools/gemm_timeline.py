#!/usr/bin/env python
"""Timeline of CTA 0 of one tcgen05 GEMM launch (xggm_debug_timeline): where a persistent CTA spends its time.
usage: python tools/gemm_timeline.py [M N K] [--mode fp32|bf16] [--resid]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xggm_b200 as X  # noqa: E402
from xggm_b200 import _lib  # noqa: E402
from xggm_b200._lib import call, ptr  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("shape", nargs="*", type=int, default=[9216, 768, 768])
ap.add_argument("--mode", default="fp32")
ap.add_argument("--resid", action="store_true")
args = ap.parse_args()
M, N, K = args.shape
dev = torch.device("cuda")
X.set_precision(args.mode)
a = torch.randn(M, K, device=dev)
w = torch.randn(N, K, device=dev) * 0.03
bias = torch.randn(N, device=dev)
resid = torch.randn(M, N, device=dev) if args.resid else None
out = torch.empty(M, N, device=dev)
work = torch.empty(_lib.load().xggm_linear_work_bytes(M, N, K), device=dev, dtype=torch.uint8)
dbg = torch.zeros(32, dtype=torch.int64, device=dev)
lib = _lib.load()
for it in range(3):
    call("xggm_linear_fwd", ptr(a), ptr(w), ptr(bias), ptr(resid), ptr(out), M, N, K, ptr(work))
torch.cuda.synchronize()
lib.xggm_debug_timeline(dbg.data_ptr())
call("xggm_linear_fwd", ptr(a), ptr(w), ptr(bias), ptr(resid), ptr(out), M, N, K, ptr(work))
torch.cuda.synchronize()
lib.xggm_debug_timeline(None)
t = dbg.cpu().tolist()
names = {0: "kernel start", 1: "setup done (barriers, TMEM)", 18: "kernel end"}
for i in range(4):
    names[2 + 4 * i] = f"tile {i}: first operands landed (MMA starts)"
    names[3 + 4 * i] = f"tile {i}: last MMA issued"
    names[4 + 4 * i] = f"tile {i}: accumulator ready (epilogue starts)"
    names[5 + 4 * i] = f"tile {i}: epilogue done"
t0 = t[0]
for slot, ts in sorted(((s, v) for s, v in enumerate(t) if v), key=lambda kv: kv[1]):
    print(f"{(ts - t0) / 1e3:9.2f} us  {names.get(slot, slot)}")
