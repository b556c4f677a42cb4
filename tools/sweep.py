#!/usr/bin/env python
"""BASELINE configs[3]: the graph block isolated, B = 256..4096 x nodes 36 / 64 / 100 x {fp32, bf16} (+ the GIN
generator and the relation branch at the BASELINE batch), one bench.py --quick line per point -> JSONL.
usage: python tools/sweep.py > gpurun_out/sweep.jsonl   (one B200; ~10 minutes)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
points = [(B, N, prec, "GCN", "node") for N in (36, 64, 100) for B in (256, 512, 1024, 2048, 4096) for prec in ("fp32", "bf16")]
points += [(256, 36, prec, gnn, br) for prec in ("fp32", "bf16") for gnn, br in (("GIN", "node"), ("GCN", "relation"), ("GIN", "relation"))]
only = sys.argv[1:]
for B, N, prec, gnn, branch in points:
    if only and str(B) not in only and f"N{N}" not in only:
        continue
    steps = max(3, min(20, 8192 // B))
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--quick", "--batch", str(B), "--nodes", str(N), "--precision", prec,
           "--gnn", gnn, "--branch", branch, "--steps", str(steps), "--warmup", "3"]
    env = dict(os.environ, XGGM_BENCH_NO_CLOCKS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if r.returncode != 0 or not lines:
        print(json.dumps({"B": B, "N": N, "precision": prec, "gnn": gnn, "branch": branch, "error": r.stderr[-300:]}), flush=True)
        continue
    d = json.loads(lines[-1])
    print(json.dumps({"B": B, "N": N, "precision": prec, "gnn": gnn, "branch": branch, "ms_per_step": d["ms_per_step"],
                      "samples_per_s": d["value"], "e2e_samples_per_s": d["e2e"]["value"],
                      "gemm_frac_algorithmic": d["roofline"]["frac"], "gemm_frac_executed": d["roofline"]["executed_frac"],
                      "gemm_share_of_step": d["roofline"]["gemm_share_of_step"],
                      "block_frac_of_bf16_peak": d["block_roofline"]["frac_of_bf16_peak"], "gpu_launches": d["gpu_launches"],
                      "steps": d["steps"]}), flush=True)
