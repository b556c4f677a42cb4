#!/usr/bin/env python
"""Micro-benchmark of the projection engine through the C ABI (xggm_linear_fwd / bwd_input / bwd_weight).
usage: python tools/gemm_bench.py [M N K] [--iters 20]   (env XGGM_TC_PAIR=0 disables the CTA-pair kernel)"""
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
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--modes", default="fp32,bf16")
ap.add_argument("--ops", default="", help="substring filter on the op names, e.g. 'fwd  bias  '")
args = ap.parse_args()
M, N, K = args.shape
dev = torch.device("cuda")
a = torch.randn(M, K, device=dev)
w = torch.randn(N, K, device=dev) * 0.03
g = torch.randn(M, N, device=dev)
bias = torch.randn(N, device=dev)
resid = torch.randn(M, N, device=dev)
out = torch.empty(M, N, device=dev)
ga = torch.empty(M, K, device=dev)
gw = torch.empty(N, K, device=dev)
work = torch.empty(_lib.load().xggm_linear_work_bytes(M, N, K), device=dev, dtype=torch.uint8)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)


def timed(fn):
    for _ in range(3):
        fn()
    _lib.gemm_profile(True)
    for _ in range(args.iters):
        flush.fill_(1)
        fn()
    torch.cuda.synchronize()
    ms, n, fl = _lib.gemm_profile()
    _lib.gemm_profile(False)
    return ms / n * 1e3, fl / (ms * 1e-3) / 1e12


for mode in args.modes.split(","):
    X.set_precision(mode)
    passes = 3 if mode == "fp32" else 1
    for name, fn in [
        ("fwd  bias      ", lambda: call("xggm_linear_fwd", ptr(a), ptr(w), ptr(bias), None, ptr(out), M, N, K, ptr(work))),
        ("fwd  bias+resid", lambda: call("xggm_linear_fwd", ptr(a), ptr(w), ptr(bias), ptr(resid), ptr(out), M, N, K, ptr(work))),
        ("dgrad          ", lambda: call("xggm_linear_bwd_input", ptr(g), ptr(w), ptr(ga), M, N, K, 0, ptr(work))),
        ("dgrad accum    ", lambda: call("xggm_linear_bwd_input", ptr(g), ptr(w), ptr(ga), M, N, K, 1, ptr(work))),
        ("wgrad          ", lambda: call("xggm_linear_bwd_weight", ptr(g), ptr(a), ptr(gw), None, M, N, K, 0, ptr(work))),
    ]:
        if args.ops and args.ops not in name:
            continue
        us, tf = timed(fn)
        print(f"{mode:5s} {name} M={M} N={N} K={K}: {us:8.1f} us/launch  {tf:7.1f} TFLOP/s algorithmic  "
              f"{tf * passes:7.1f} executed ({tf * passes / 1601.2:.2f} of measured bf16 peak)", flush=True)
X.set_precision("fp32")
