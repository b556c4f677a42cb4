"""One full X-GGM training ITERATION (VQA-CP v2 recipe) around the library's graph block -- measurement harness for
`bench.py --workload iteration` (BASELINE configs[1] fp32 / configs[2] bf16).

An iteration is the body of the reference trainer's loop (src/vqa/vqacpv2.py:164-254), two optimiser steps:

  STEP A  plain VQA     logit = logit_fc(LXMERT(feats, boxes, sent).pooled); BCE * A; backward; clip 5; BertAdam   :170-177
  STEP B  GGM           LXMERT forward again (weights just changed), node branch (--delta 0, the shipped recipe;
                        `delta` > 0 mixes in the relation branch through a rank-synchronous BranchSchedule),
                        logit_fc(x_gen); BCE * A + 1.1 * loss_sm; backward; clip 5; BertAdam                        :183-254

What runs where: the LXMERT encoder is STOCK PyTorch (tools/lxmert_torch.py: out of the hot path, shared by every arm,
optionally under bf16 autocast); the graph block, the answer head, BCE, clip_grad_norm_ and BertAdam are the
library's kernels (xggm_b200).  All gradients live in two flat buckets (encoder 207.9 M, down-task 12.9 M
parameters -- the reference's two learning-rate groups, src/vqa/vqacpv2.py:113-128) that NCCL averages once per
optimiser step: 2 x 0.88 GB per iteration at fp32, the volume SURVEY 8e names.

SURVEY 8 f-2: the batch is copied to the device ONCE per iteration (the reference calls .cuda() on feats / boxes in
both steps, :171 and :185) and the copy of iteration i+1 runs on a side stream while iteration i computes.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import lxmert_torch as LT  # noqa: E402

N_OBJ, FEAT, HID, NUM_ANS, T_LANG = 36, 2048, 768, 2274, 20


def synthetic_batch(seed, B, num_answers=NUM_ANS):
    """SURVEY 8d synthetic inputs of one VQA-CP v2 batch, as pinned host tensors."""
    g = torch.Generator().manual_seed(seed)
    feats = torch.relu(torch.randn(B, N_OBJ, FEAT, generator=g))
    xy = torch.rand(B, N_OBJ, 2, 2, generator=g).sort(dim=-1)[0]
    boxes = torch.stack([xy[..., 0, 0], xy[..., 1, 0], xy[..., 0, 1], xy[..., 1, 1]], dim=-1)
    ids, mask = LT.synthetic_language(B, T_LANG, 30522, seed + 1)
    target = torch.zeros(B, num_answers)
    idx = torch.randint(0, num_answers, (B, 3), generator=g)
    val = torch.tensor([0.3, 0.6, 0.9, 1.0])[torch.randint(0, 4, (B, 3), generator=g)]
    target.scatter_(1, idx, val)
    c = torch.rand(B, N_OBJ, N_OBJ, generator=g) * 1.2 - 0.2
    c = c + c.transpose(1, 2)
    adj = c / c.amax(dim=(1, 2), keepdim=True)
    out = [feats, boxes, ids, mask, target, adj]
    if torch.cuda.is_available():
        out = [t.pin_memory() for t in out]
    return out


class XGGMIteration:
    """The library arm: stock LXMERT + xggm_b200 block / heads / optimiser.  `autocast`: run the stock encoder under
    torch.autocast(bf16) and the block on the single-pass bf16 engine (BASELINE configs[2])."""

    def __init__(self, dev, B, autocast=False, lr=1e-6, t_total=100000, delta=0, num_answers=NUM_ANS, library_visn_fc=True,
                 overlap=True):
        import xggm_b200 as X
        from xggm_b200.ddp import BranchSchedule, FlatGrads
        self.X, self.dev, self.B, self.autocast, self.A = X, dev, B, autocast, num_answers
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        X.set_precision("bf16" if autocast else "fp32")
        torch.manual_seed(9595)
        self.lxmert = LT.LXRTFeatureExtractionTorch(LT.Config(), X.VisualFeatEncoder if library_visn_fc else None).to(dev).train()
        self.heads = X.XGGMHeads(HID, "GCN", 2, N_OBJ).to(dev).train()
        self.answer = X.AnswerHead(HID, num_answers).to(dev).train()
        base = list(self.lxmert.parameters())
        down = list(self.answer.parameters()) + list(self.heads.parameters())
        # ONE bucket for both learning-rate groups (encoder lr, down-task 4 lr; vqacpv2.py:113-128; t_total counts both
        # optimiser steps of an iteration): at N > 1 it lives in symmetric memory and the all-reduce + joint clip +
        # BertAdam + parameter all-gather run as the fused peer-memory step (XGGM_DP_FUSED=0: NCCL all-reduce instead)
        want_fused = self.world > 1 and os.environ.get("XGGM_DP_FUSED", "1") != "0"
        self.fg = FlatGrads(base + down, symmetric=want_fused)
        self.fg_base = self.fg_down = self.fg          # (names kept for callers that size the buckets)
        self.optim = X.BertAdam([{"params": base, "lr": lr}, {"params": down, "lr": 4 * lr}], lr=lr, warmup=0.1,
                                t_total=t_total, flat_grads=self.fg)
        self.fused_dp = want_fused and self.optim.fused_allreduce_available()
        self.branch = BranchSchedule(delta)
        self.overlap = overlap and self.world > 1
        self.side = torch.cuda.Stream(device=dev)
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.staged, self.stage_free = torch.cuda.Event(), torch.cuda.Event()
        self.stage = None
        self.cur = None

    # ---- input staging (one H2D per iteration, overlapped with the previous iteration) ----------------------
    def prefetch(self, host_batch):
        if self.stage is None:
            self.stage = [torch.empty_like(t, device=self.dev) for t in host_batch]
            self.cur = [torch.empty_like(t, device=self.dev) for t in host_batch]
            self.stage_free.record(torch.cuda.current_stream(self.dev))
        self.copy_stream.wait_event(self.stage_free)
        with torch.cuda.stream(self.copy_stream):
            for s, t in zip(self.stage, host_batch):
                s.copy_(t, non_blocking=True)
            self.staged.record(self.copy_stream)

    def take(self):
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_event(self.staged)
        for d, s in zip(self.cur, self.stage):
            d.copy_(s, non_blocking=True)
        self.stage_free.record(cur)
        return self.cur

    def h2d_bytes(self, host_batch):
        return sum(t.numel() * t.element_size() for t in host_batch)

    # ---- the two optimiser steps ---------------------------------------------------------------------------
    def _encode(self, feats, boxes, ids, mask):
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.autocast):
            (lang, visn), pooled = self.lxmert(ids, None, mask, visual_feats=(feats, boxes))
        return visn.float(), pooled.float()

    def _reduce_and_step(self):
        X = self.X
        if self.fused_dp:
            self.optim.step_allreduce(5.0)
            return
        if self.world > 1:
            self.fg.all_reduce(average=True)
        self.optim.step(X.clip_grad_norm_(self.fg, 5.0))

    def step_a(self, feats, boxes, ids, mask, target):
        X = self.X
        self.optim.zero_grad()
        _, pooled = self._encode(feats, boxes, ids, mask)
        logit = self.answer(pooled)
        loss = X.bce_with_logits(logit, target, scale=self.A)
        loss.backward()
        self._reduce_and_step()
        return loss.detach()

    def step_b(self, feats, boxes, ids, mask, target, adj_true, branch=None):
        X = self.X
        self.optim.zero_grad()
        visn, pooled = self._encode(feats, boxes, ids, mask)
        if (branch or self.branch.next()) == "relation":
            x_gen, loss_sm, _, _ = self.heads.relation_step(pooled, visn, adj_true, 1.0, self.A, kl_weight=8.0)
            w = 6.0
        else:
            x_gen, loss_sm, _, _ = self.heads.node_step(pooled, visn, adj_true, 1.0, self.A)
            w = 1.1
        logit = self.answer(x_gen)
        loss = X.bce_with_logits(logit, target, scale=self.A) + w * loss_sm
        loss.backward()
        self._reduce_and_step()
        return loss.detach()

    def iteration(self, batch, branch=None):
        """branch: None = draw it from the rank-synchronous schedule (eager runs); 'node' / 'relation' = fixed (one
        CUDA graph is captured per branch and the schedule picks the graph to replay)."""
        feats, boxes, ids, mask, target, adj = batch
        self.step_a(feats, boxes, ids, mask, target)
        return self.step_b(feats, boxes, ids, mask, target, adj, branch)

    def block_only(self, batch):
        """Just the library's part of step B on fixed encoder outputs (for the share-of-iteration figure)."""
        feats, boxes, ids, mask, target, adj = batch
        with torch.no_grad():
            visn, pooled = self._encode(feats, boxes, ids, mask)
        return visn, pooled
