#!/usr/bin/env python
"""One eager forward+backward step of the bench workload between cudaProfilerStart/Stop, for
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/step_for_ncu.py
Flags mirror bench.py (--batch/--nodes/--precision/--gnn)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import xggm_b200 as X  # noqa: E402
from bench import synthetic_inputs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--nodes", type=int, default=36)
ap.add_argument("--precision", default="fp32")
ap.add_argument("--gnn", default="GCN")
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--branch", default="node", choices=["node", "relation"])
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(9595)
X.set_precision(a.precision)
model = X.XGGMHeads(768, a.gnn, 2, a.nodes).to(dev).train()
from xggm_b200.ddp import FlatGrads  # noqa: E402
grads = FlatGrads(model.parameters())
optim = X.BertAdam(model.parameters(), lr=4e-6, flat_grads=grads)
visn, xp, adj = (t.to(dev) for t in synthetic_inputs(9596, a.batch, a.nodes, 768))
cot = torch.randn(a.batch, 768, device=dev)
w_rel, w_node = torch.tensor(6.0, device=dev), torch.tensor(1.1, device=dev)


def compute():
    grads.zero_()
    x = xp.detach().requires_grad_(True)
    feat = visn.detach().requires_grad_(True)
    if a.branch == "relation":
        x_gen, loss_sm, _, _ = model.relation_step(x, feat, adj, 1.0, 2274, kl_weight=12.0)
        torch.autograd.backward([x_gen, loss_sm], [cot, w_rel])
    else:
        x_gen, loss_sm, _, _ = model.node_step(x, feat, adj, 1.0, 2274)
        torch.autograd.backward([x_gen, loss_sm], [cot, w_node])
    optim.step(X.clip_grad_norm_(grads, 5.0))


for _ in range(3):
    compute()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(a.steps):
    compute()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
