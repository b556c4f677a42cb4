"""GGM part of the reference's model containers and the two GGM training branches.

``XGGMHeads`` owns exactly the sub-modules ``VQAModel`` / ``GQAModel`` add on top of the
LXMERT encoder for graph generative modeling -- ``generator``, ``encoder_adj``,
``node_fc``, ``fusion_fc`` (src/vqa/vqacpv2_model.py:71-105, src/gqa/gqa_ood_model.py:71-112)
-- under the same attribute names, so their ``state_dict`` keys match a reference
checkpoint's.  ``relation_step`` / ``node_step`` are the bodies of the two branches of the
trainers' GGM step (src/vqa/vqacpv2.py:194-218 and :226-247; src/gqa/gqa_ood.py:178-202 and
:236-252) from the LXMERT outputs to ``(x_gen, loss_sm)``.
"""
import torch
import torch.nn as nn

from . import functional as XF
from . import glue
from .nn import GATGenerator, GCNGenerator, GeLU, GINGenerator

GENERATORS = {"GCN": GCNGenerator, "GIN": GINGenerator, "GAT": GATGenerator}


class AnswerHead(nn.Module):
    """``logit_fc`` of VQAModel / GQAModel (src/vqa/vqacpv2_model.py:63-69):
    Linear(H, 2H) -> GeLU -> BertLayerNorm(2H, eps 1e-12) -> Linear(2H, num_answers), initialised as
    ``init_bert_weights`` does (N(0, 0.02) weights, zero biases, unit LayerNorm; src/lxrt/modeling.py:734-747).
    Sub-module indices match the reference's nn.Sequential, so ``logit_fc.*`` checkpoint keys load unchanged."""

    def __init__(self, hid_dim=768, num_answers=2274, initializer_range=0.02):
        super().__init__()
        self.logit_fc = nn.Sequential(nn.Linear(hid_dim, hid_dim * 2), GeLU(), nn.LayerNorm(hid_dim * 2, eps=1e-12),
                                      nn.Linear(hid_dim * 2, num_answers))
        for m in self.logit_fc:
            if isinstance(m, nn.Linear):
                m.weight.data.normal_(mean=0.0, std=initializer_range)
                m.bias.data.zero_()

    def forward(self, x):
        m = self.logit_fc
        z = XF.linear(x, m[0].weight, m[0].bias)
        h = XF.gelu_ln_drop(z, m[2].weight, m[2].bias, None, 0.0, m[2].eps)
        return XF.linear(h, m[3].weight, m[3].bias)


class XGGMHeads(nn.Module):
    def __init__(self, hid_dim=768, gnn="GCN", n_layers=2, n_nodes=36):
        super().__init__()
        if gnn not in GENERATORS:
            raise ModuleNotFoundError(gnn)  # as src/vqa/vqacpv2_model.py:85-86
        self.hid_dim, self.n_nodes = hid_dim, n_nodes
        self.generator = GENERATORS[gnn](hidden_dim=hid_dim, n_layers=n_layers)
        # relation / node initialisation heads, src/vqa/vqacpv2_model.py:91-105
        self.encoder_adj = nn.Sequential(nn.Linear(hid_dim, n_nodes * (n_nodes - 1) // 2), nn.Sigmoid())
        self.node_fc = nn.Sequential(nn.Linear(hid_dim, hid_dim), GeLU(), nn.LayerNorm(hid_dim))
        self.fusion_fc = nn.Sequential(nn.Linear(hid_dim * 2, hid_dim), GeLU(), nn.LayerNorm(hid_dim))

    # -- pieces ---------------------------------------------------------------
    def init_relation(self, x):
        """encoder_adj(x) scattered to a symmetric zero-diagonal adjacency (vqacpv2.py:195-199)."""
        lin = self.encoder_adj[0]
        v = XF.sigmoid(XF.linear(x, lin.weight, lin.bias))
        return glue.triu_scatter(v, self.n_nodes)

    def init_nodes(self, x):
        """node_fc applied once per sample; the 36-fold repeat (vqacpv2.py:228-229) is folded
        into the noise kernel's broadcast read, its backward into a node sum."""
        m = self.node_fc
        z = XF.linear(x, m[0].weight, m[0].bias)
        return XF.gelu_ln_drop(z, m[2].weight, m[2].bias, None, 0.0, m[2].eps)

    def fuse(self, x, node_feats):
        """fusion_fc(cat[x, tanh(mean_n node_feats)]) (vqacpv2.py:216-218)."""
        m = self.fusion_fc
        cat = glue.fuse_readout(x, node_feats)
        z = XF.linear(cat, m[0].weight, m[0].bias)
        return XF.gelu_ln_drop(z, m[2].weight, m[2].bias, None, 0.0, m[2].eps)

    # -- the two GGM branches ----------------------------------------------------
    def relation_step(self, x, feat, adj_true, sigma, num_answers, kl_weight=8.0, randn=None):
        """x [B,H] pooled, feat [B,N,H] = feat_seq[1], adj_true [B,N,N] raw obj36_adj.
        kl_weight: 8 for VQA-CP v2 (vqacpv2.py:212), 12 for GQA-OOD (gqa_ood.py:197)."""
        adj_t = glue.strip_diag(adj_true)
        adj_noise = self.init_relation(x)
        adj_noise, grad_log_noise = glue.add_edge_noise(adj_noise, sigma=sigma, randn=randn)
        node_feats, adj_gen = self.generator(feat, adj_noise)
        loss_grad = glue.loss_func(adj_gen, grad_log_noise, sigma=sigma)
        d_loss = glue.compute_kl_loss(adj_t, adj_gen) * num_answers
        loss_sm = kl_weight * d_loss + loss_grad
        return self.fuse(x, node_feats), loss_sm, node_feats, adj_gen

    def node_step(self, x, feat, adj_true, sigma, num_answers, randn=None):
        adj_t = glue.strip_diag(adj_true)
        nodes0 = self.init_nodes(x)
        node_feats, feat_grad = glue.add_feature_noise(nodes0, sigma=sigma, randn=randn, n_nodes=self.n_nodes)
        node_feats, adj_gen = self.generator(node_feats, adj_t)
        # loss_sm = 0.15 * compute_kl_loss(node_feats, feat) * A + 6 * loss_func(node_feats, feat_grad, sigma)
        # and the fusion_fc input, one pass over node_feats per direction (xggm_node_tail_*)
        loss_sm, cat = XF.node_tail(node_feats, feat, feat_grad, x, sigma, 0.15 * num_answers, 6.0)
        m = self.fusion_fc
        z = XF.linear(cat, m[0].weight, m[0].bias)
        x_gen = XF.gelu_ln_drop(z, m[2].weight, m[2].bias, None, 0.0, m[2].eps)
        return x_gen, loss_sm, node_feats, adj_gen
