#!/usr/bin/env python
"""Micro-benchmark of the non-GEMM kernels at the BASELINE shape (B=256, N=36, H=768) through the C ABI.
usage: python tools/op_bench.py [--ops adj,ln,gld] [--iters 20]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xggm_b200 as X  # noqa: E402
import xggm_b200.functional as XF  # noqa: E402
from xggm_b200._lib import call, ptr  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ops", default="adj,ln,gld,regen,visual")
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("-B", type=int, default=256)
args = ap.parse_args()
B, N, H = args.B, 36, 768
dev = torch.device("cuda")
x = torch.randn(B, N, H, device=dev)
g = torch.randn(B, N, H, device=dev)
adj = torch.rand(B, N, N, device=dev)
out = torch.empty_like(x)
gamma, beta = torch.ones(H, device=dev), torch.zeros(H, device=dev)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
M = B * N
T = M * H * 4 / 1e6  # MB per [M,H] fp32 tensor


def timed(name, fn, mbytes):
    for _ in range(3):
        fn()
    evs = []
    for _ in range(args.iters):
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        evs.append((s, e))
    torch.cuda.synchronize()
    us = sorted(s.elapsed_time(e) * 1e3 for s, e in evs)[len(evs) // 2]
    print(f"{name:34s} {us:8.1f} us   {mbytes:7.1f} MB algorithmic -> {mbytes / us:6.2f} TB/s", flush=True)


ops = args.ops.split(",")
if "adj" in ops:
    gx, graw = torch.empty_like(x), torch.empty_like(adj)
    for tag, wk in (("SIMT", None), ("tensor cores", XF._adj_work(B, N, H, dev))):
        timed(f"adj_apply fwd [{tag}]", lambda: call("xggm_adj_apply_fwd", ptr(adj), ptr(x), ptr(out), B, N, H, 1.0, None, 0.0, ptr(wk)), 2 * T)
        timed(f"adj_apply bwd adj^T g [{tag}]", lambda: call("xggm_adj_apply_bwd", ptr(adj), ptr(x), ptr(g), ptr(gx), None, B, N, H, 1.0, None, 0.0, 0, ptr(wk)), 2 * T)
        timed(f"adj_apply bwd accumulate [{tag}]", lambda: call("xggm_adj_apply_bwd", ptr(adj), ptr(x), ptr(g), ptr(gx), None, B, N, H, 1.0, None, 0.0, 1, ptr(wk)), 3 * T)
if "ln" in ops:
    h, xhat, rstd = torch.empty_like(x), torch.empty_like(x), torch.empty(M, device=dev)
    timed("layernorm fwd (h, xhat)", lambda: call("xggm_layernorm_fwd", ptr(x), ptr(gamma), ptr(beta), ptr(h), ptr(xhat), ptr(rstd), M, H, 1e-5), 3 * T)
    gu, gg, gb = torch.empty_like(x), torch.zeros(H, device=dev), torch.zeros(H, device=dev)
    timed("layernorm bwd", lambda: call("xggm_layernorm_bwd", ptr(g), ptr(xhat), ptr(rstd), ptr(gamma), ptr(gu), ptr(gg), ptr(gb), M, H), 3 * T)
if "gld" in ops:
    keep = (torch.rand(M, H, device=dev) > 0.5).to(torch.uint8)
    mean, rstd2 = torch.empty(M, device=dev), torch.empty(M, device=dev)
    timed("gelu_ln_drop fwd (mask, accumulate)", lambda: call("xggm_gelu_ln_drop_fwd", ptr(x), ptr(gamma), ptr(beta), ptr(keep), 2.0, ptr(out), ptr(mean), ptr(rstd2), M, H, 1e-5, 1), 3.25 * T)
    gz, gg, gb = torch.empty_like(x), torch.zeros(H, device=dev), torch.zeros(H, device=dev)
    timed("gelu_ln_drop bwd (mask)", lambda: call("xggm_gelu_ln_drop_bwd", ptr(g), ptr(x), ptr(mean), ptr(rstd2), ptr(gamma), ptr(keep), 2.0, ptr(gz), ptr(gg), ptr(gb), M, H), 3.25 * T)
if "visual" in ops:
    enc = X.VisualFeatEncoder().to(dev).train()
    feats = torch.relu(torch.randn(B, N, 2048, device=dev))
    boxes = torch.rand(B, N, 4, device=dev)

    def vfe():
        f = feats.requires_grad_(True)
        o = enc((f, boxes))
        o.backward(g)
        f.grad = None
    timed("VisualFeatEncoder fwd+bwd (eager)", vfe, 2 * M * 2048 * 4 / 1e6 + 6 * T)
if "regen" in ops:
    timed("adj_regen fwd (tensor-core Gram)", lambda: XF.adj_regen(x), T)
