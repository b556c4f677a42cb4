"""One GGM training step of the reference path on any torch device -- BASELINE INFRASTRUCTURE, NOT PRODUCT.

Used by bench.py's `cpu_baseline` / `--impl reference` legs (device "cpu") and by its `eager_gpu_baseline` leg
(device "cuda": the reference's eager PyTorch path on the same B200, SURVEY section 2a / BASELINE.md section 1).

kind "reference": the UNMODIFIED reference modules vendored into oracle/_ref/ by oracle/vendor_ref.py
    (module.graph_generative_modeling.GCNGenerator / GINGenerator, module.graph_utils.add_*_noise_v2,
    lxrt.modeling.GeLU, lxrt.optimization.BertAdam, loss_func / compute_kl_loss cut out of vqa/vqacpv2.py with
    `ast`), wired by the trainer's own glue lines (src/vqa/vqacpv2.py:187-254, restated below with line numbers:
    they have no callable entry point).
kind "port": oracle/xggm_oracle.py (when oracle/_ref is absent).

The step is what bench.py's GPU arm times: GGM branch forward, backward to every parameter and to the LXMERT
outputs, clip_grad_norm_(5.), BertAdam.step.  Dropout and noise are drawn by torch as the reference does.
"""
import ast
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref", "src")


def reference_available():
    from oracle import vendor_ref
    return vendor_ref.vendored()


_ref_cache = None


def _import_reference():
    global _ref_cache
    if _ref_cache is not None:
        return _ref_cache
    for m in ("boto3", "botocore", "botocore.exceptions"):      # download helpers of lxrt/file_utils.py: never called
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules["botocore.exceptions"].ClientError = Exception
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import module.graph_generative_modeling as ggm
    import module.graph_utils as gu
    from lxrt.modeling import GeLU
    from lxrt.optimization import BertAdam
    tree = ast.parse(open(os.path.join(REF, "vqa", "vqacpv2.py")).read())
    ns = {"torch": torch, "F": torch.nn.functional}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("loss_func", "compute_kl_loss"):
            exec(compile(ast.Module([node], []), "vqacpv2.py", "exec"), ns)
    _ref_cache = (ggm, gu, GeLU, BertAdam, ns["loss_func"], ns["compute_kl_loss"])
    return _ref_cache


def synthetic_inputs(seed, B, n_nodes=36, hidden=768):
    from oracle import xggm_oracle as O
    return O.make_inputs(seed, B, n_nodes, hidden)


class ReferenceStep:
    """step() runs one iteration; .kind says which implementation is underneath."""

    def __init__(self, device, B, branch="node", gnn="GCN", hidden=768, n_layers=2, n_nodes=36, sigma=1.0,
                 num_answers=2274, lr=4e-6, seed=9595, prefer_reference=True):
        self.device = torch.device(device)
        self.B, self.branch, self.gnn, self.H, self.L, self.N = B, branch, gnn, hidden, n_layers, n_nodes
        self.sigma, self.A, self.lr = sigma, num_answers, lr
        visn, xp, adj_true = synthetic_inputs(seed + 1, B, n_nodes, hidden)
        self.visn, self.xp, self.adj_true = (t.to(self.device) for t in (visn, xp, adj_true))
        self.cot = torch.randn(B, hidden, generator=torch.Generator().manual_seed(1)).to(self.device)
        self.kind = "reference" if (prefer_reference and reference_available()) else "port"
        torch.manual_seed(seed)
        if self.kind == "reference":
            self._build_reference()
        else:
            self._build_port(seed)

    # ------------------------------------------------------------------ the real reference modules
    def _build_reference(self):
        import torch.nn as nn
        ggm, gu, GeLU, BertAdam, loss_func, compute_kl_loss = _import_reference()
        H, N = self.H, self.N
        cls = {"GCN": ggm.GCNGenerator, "GIN": ggm.GINGenerator}[self.gnn]
        self.generator = cls(hidden_dim=H, n_layers=self.L)                                # vqacpv2_model.py:71-85
        self.encoder_adj = nn.Sequential(nn.Linear(H, N * (N - 1) // 2), nn.Sigmoid())     # :91-94
        self.node_fc = nn.Sequential(nn.Linear(H, H), GeLU(), nn.LayerNorm(H))             # :95-99
        self.fusion_fc = nn.Sequential(nn.Linear(H * 2, H), GeLU(), nn.LayerNorm(H))       # :101-105
        self.mods = [self.generator, self.encoder_adj, self.node_fc, self.fusion_fc]
        for m in self.mods:
            m.to(self.device).train()
        self.params = [p for m in self.mods for p in m.parameters()]
        self.optim = BertAdam(self.params, lr=self.lr, warmup=-1, t_total=-1)              # constant lr, as the GPU arm
        self.gu, self.loss_func, self.compute_kl_loss = gu, loss_func, compute_kl_loss

    def _step_reference(self):
        gu, loss_func, compute_kl_loss = self.gu, self.loss_func, self.compute_kl_loss
        for p in self.params:
            p.grad = None                                                                   # model.zero_grad(), :170
        x = self.xp.clone().requires_grad_(True)              # pooled LXMERT output
        feat = self.visn.clone().requires_grad_(True)         # feat_seq[1]
        adj_true = self.adj_true.triu(1) + self.adj_true.tril(-1)                           # :188
        if self.branch == "relation":
            adj_noise = torch.zeros_like(adj_true)                                          # :196
            adj_temp = torch.ones_like(adj_true).triu(1)                                    # :197
            adj_noise[adj_temp == 1] = self.encoder_adj(x).view(-1)                         # :198
            adj_noise = adj_noise + adj_noise.transpose(1, 2)                               # :199
            adj_noise, grad_log_noise = gu.add_edge_noise_v2(adj_noise, sigma=self.sigma)   # :201
            node_feats, adj_gen = self.generator(feat, adj_noise)                           # :204
            loss_grad = loss_func(adj_gen, grad_log_noise, sigma=self.sigma)                # :208
            d_loss = compute_kl_loss(adj_true, adj_gen) * self.A                            # :210
            loss_sm = 12 * d_loss + loss_grad                                               # gqa_ood.py:197
            w = 6.0                                                                         # :221
        else:
            node_feats = x.unsqueeze(1).repeat(1, self.N, 1)                                # :228
            node_feats = self.node_fc(node_feats)                                           # :229
            node_feats, feat_grad = gu.add_feature_noise_v2(node_feats, sigma=self.sigma)   # :230
            node_feats, _ = self.generator(node_feats, adj_true)                            # :233
            d_loss = compute_kl_loss(node_feats, feat) * self.A                             # :237
            loss_grad = loss_func(node_feats, feat_grad, sigma=self.sigma)                  # :239
            loss_sm = 0.15 * d_loss + 6 * loss_grad                                         # :241
            w = 1.1                                                                         # :250
        x_gen = self.fusion_fc(torch.cat([x, torch.tanh(node_feats.mean(1))], dim=-1))      # :216-218 / :245-246
        loss = (x_gen * self.cot).sum() + w * loss_sm         # the answer head's place: a fixed cotangent on x_gen
        loss.backward()                                                                     # :251
        torch.nn.utils.clip_grad_norm_(self.params, 5.)                                     # :252
        self.optim.step()                                                                   # :253
        return loss_sm.detach()

    # ------------------------------------------------------------------ the oracle port
    def _build_port(self, seed):
        from oracle import xggm_oracle as O
        self.O = O
        self.p = {k: v.to(self.device).requires_grad_(True)
                  for k, v in O.make_params(seed, self.gnn, self.H, self.L, self.N, heads=True).items()}
        self.mom = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in self.p.items()}

    def _step_port(self):
        O, p, B, N, H = self.O, self.p, self.B, self.N, self.H
        x = self.xp.clone().requires_grad_(True)
        feat = self.visn.clone().requires_grad_(True)
        nh = 3 if self.gnn == "GCN" else 2
        keeps = [[(torch.rand(B, N, H, device=self.device) >= 0.5) for _ in range(nh)] for _ in range(self.L)]
        if self.branch == "relation":
            randn = torch.randn(B, N, N, device=self.device)
            x_gen, loss_sm, _, _ = O.relation_branch(x, feat, self.adj_true, p, self.sigma, randn, keeps, self.A, self.gnn,
                                                     self.L, 12.0)
            w = 6.0
        else:
            randn = torch.randn(B, N, H, device=self.device)
            x_gen, loss_sm, _, _ = O.node_branch(x, feat, self.adj_true, p, self.sigma, randn, keeps, self.A, self.gnn, self.L)
            w = 1.1
        ((x_gen * self.cot).sum() + w * loss_sm).backward()
        live = [k for k, v in p.items() if v.grad is not None]
        total = torch.sqrt(sum((p[k].grad.double() ** 2).sum() for k in live))
        coef = torch.clamp(5.0 / (total + 1e-6), max=1.0).float()
        with torch.no_grad():
            for k in live:
                new_p, m, v2 = O.bertadam_step(p[k], p[k].grad * coef, self.mom[k][0], self.mom[k][1], self.lr)
                p[k].copy_(new_p)
                self.mom[k] = (m, v2)
        for v in p.values():
            v.grad = None
        return loss_sm.detach()

    def step(self):
        return self._step_reference() if self.kind == "reference" else self._step_port()


class ReferenceIteration:
    """One full trainer iteration (src/vqa/vqacpv2.py:164-254: step A plain VQA, step B GGM node branch) built from
    the UNMODIFIED reference modules vendored in oracle/_ref -- LXRTFeatureExtraction (9/5/5 layers, stock PyTorch),
    logit_fc, GCNGenerator + heads, the reference's BertAdam with its two learning-rate groups -- for the
    eager-PyTorch-on-B200 baseline of `bench.py --workload iteration`.  Needs oracle/_ref (there is no port of the
    LXMERT encoder).  `autocast`: torch.autocast(bf16) around the model calls (the reference has no reduced-precision
    path of its own)."""

    def __init__(self, device, autocast=False, lr=1e-6, t_total=100000, num_answers=2274, hidden=768, n_nodes=36, sigma=1.0):
        import torch.nn as nn
        ggm, gu, GeLU, BertAdam, loss_func, compute_kl_loss = _import_reference()
        import lxrt.modeling as RM
        self.device, self.autocast, self.A, self.N, self.sigma = torch.device(device), autocast, num_answers, n_nodes, sigma
        RM.VISUAL_CONFIG.l_layers, RM.VISUAL_CONFIG.x_layers, RM.VISUAL_CONFIG.r_layers = 9, 5, 5   # entry.py:24-34 + param.py
        torch.manual_seed(9595)
        cfg = RM.BertConfig(vocab_size_or_config_json_file=30522)
        self.lxmert = RM.LXRTFeatureExtraction(cfg, mode="lxr")
        H = hidden
        self.logit_fc = nn.Sequential(nn.Linear(H, H * 2), GeLU(), RM.BertLayerNorm(H * 2, eps=1e-12),
                                      nn.Linear(H * 2, num_answers))                               # vqacpv2_model.py:63-69
        self.logit_fc.apply(self.lxmert.init_bert_weights)
        self.generator = ggm.GCNGenerator(hidden_dim=H, n_layers=2)
        self.node_fc = nn.Sequential(nn.Linear(H, H), GeLU(), nn.LayerNorm(H))
        self.fusion_fc = nn.Sequential(nn.Linear(H * 2, H), GeLU(), nn.LayerNorm(H))
        self.encoder_adj = nn.Sequential(nn.Linear(H, n_nodes * (n_nodes - 1) // 2), nn.Sigmoid())
        self.mods = [self.lxmert, self.logit_fc, self.generator, self.node_fc, self.fusion_fc, self.encoder_adj]
        for m in self.mods:
            m.to(self.device).train()
        base = list(self.lxmert.parameters())
        down = [p for m in self.mods[1:] for p in m.parameters()]
        self.params = base + down
        self.optim = BertAdam([{"params": down}, {"params": base, "lr": lr}], lr=4 * lr, warmup=0.1,
                              t_total=t_total)                                                      # vqacpv2.py:113-128
        self.bce = nn.BCEWithLogitsLoss()
        self.gu, self.loss_func, self.compute_kl_loss = gu, loss_func, compute_kl_loss

    def _model(self, feats, boxes, ids, mask):
        with torch.autocast(self.device.type, dtype=torch.bfloat16, enabled=self.autocast):
            (lang, visn), pooled = self.lxmert(ids, None, mask, visual_feats=(feats, boxes))
        return visn.float(), pooled.float()

    def iteration(self, host_batch):
        """host_batch: (feats, boxes, ids, mask, target, adj) on the HOST; copied to the device in both steps, as
        the reference does (vqacpv2.py:171,185)."""
        feats_h, boxes_h, ids_h, mask_h, target_h, adj_h = host_batch
        dev = self.device
        ids, mask, target = ids_h.to(dev), mask_h.to(dev), target_h.to(dev)                        # :166-167
        # ---- step A (:170-177)
        for p in self.params:
            p.grad = None
        _, x = self._model(feats_h.to(dev), boxes_h.to(dev), ids, mask)
        logit = self.logit_fc(x)
        loss = self.bce(logit, target) * logit.size(1)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.params, 5.)
        self.optim.step()
        # ---- step B, node branch (:183-254 with --delta 0)
        for p in self.params:
            p.grad = None
        feat, x = self._model(feats_h.to(dev), boxes_h.to(dev), ids, mask)                         # :185
        adj_true = adj_h.to(dev)
        adj_true = adj_true.triu(1) + adj_true.tril(-1)                                             # :187-188
        node_feats = x.unsqueeze(1).repeat(1, self.N, 1)                                            # :228
        node_feats = self.node_fc(node_feats)
        node_feats, feat_grad = self.gu.add_feature_noise_v2(node_feats, sigma=self.sigma)
        node_feats, _ = self.generator(node_feats, adj_true)
        d_loss = self.compute_kl_loss(node_feats, feat) * logit.size(1)
        loss_grad = self.loss_func(node_feats, feat_grad, sigma=self.sigma)
        loss_sm = 0.15 * d_loss + 6 * loss_grad
        x_gen = self.fusion_fc(torch.cat([x, torch.tanh(node_feats.mean(1))], dim=-1))
        logit = self.logit_fc(x_gen)
        loss = self.bce(logit, target) * logit.size(1) + 1.1 * loss_sm
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.params, 5.)
        self.optim.step()
        return loss.detach()
