"""nn.Module boundary of the X-GGM graph block.

Class names, constructor signatures, sub-module attribute names and therefore
``state_dict`` keys/shapes are those of the reference
(src/module/{gcn,gin,gat}.py, src/module/graph_generative_modeling.py), so the
LXMERT-based VQA-CP v2 / GQA-OOD containers can swap these in and reference
checkpoints load unchanged.  ``nn.Linear`` / ``nn.LayerNorm`` instances are used
as *parameter holders only* (same names, same default initialisation); every
forward pass below runs the library's CUDA kernels via ``xggm_b200.functional``.
"""
import math

import torch
import torch.nn as nn

from . import ddp
from . import functional as XF


class GeLU(nn.Module):
    """Exact-erf GeLU (reference: src/lxrt/modeling.py:127-140)."""

    def forward(self, x):
        return XF.gelu(x)


def _head(in_dim, out_dim):
    # Linear -> GeLU -> LayerNorm, indices 0/1/2 as in src/module/gcn.py:44-47
    return nn.Sequential(nn.Linear(in_dim, out_dim), GeLU(), nn.LayerNorm(out_dim))


def _head_params(seq):
    return [seq[0].weight, seq[0].bias, seq[2].weight, seq[2].bias]


def _check_uniform(input_dim, hidden_dims, n_layers, who):
    dims = [input_dim] + list(hidden_dims[:max(n_layers, 1)]) + [hidden_dims[-2], hidden_dims[-1]]
    if any(d != input_dim for d in dims):
        raise NotImplementedError(
            f"xggm_b200.{who}: only equal feature widths are supported (the reference generators "
            f"build hidden_dims=[H, H]; the residual in GCNConv requires it anyway), got {dims}")


def _head_forward(seq, h, keep=None, drop_p=0.0):
    z = XF.linear(h, seq[0].weight, seq[0].bias)
    return XF.gelu_ln_drop(z, seq[2].weight, seq[2].bias, keep, drop_p, seq[2].eps)


# ---------------------------------------------------------------------------
# GCN  (reference: src/module/gcn.py)
# ---------------------------------------------------------------------------
class GCNConv(nn.Module):
    """LN(x + dropout(W_ctx (adj @ x))) -- src/module/gcn.py:10-29."""

    def __init__(self, dim_hidden, dropout=0.0):
        super().__init__()
        self.ctx_layer = nn.Linear(dim_hidden, dim_hidden, bias=False)
        self.layer_norm = nn.LayerNorm(dim_hidden)
        self.dropout = nn.Dropout(p=dropout)

    def forward(self, x, adj):
        agg = XF.adj_apply(adj, x)
        if self.dropout.p > 0.0 and self.training:
            u = x + XF.dropout(XF.linear(agg, self.ctx_layer.weight), self.dropout.p, True)
        else:
            u = XF.linear(agg, self.ctx_layer.weight, None, x)
        return XF.layer_norm(u, self.layer_norm.weight, self.layer_norm.bias, self.layer_norm.eps)


class GCN(nn.Module):
    """Conv chain + jump-knowledge read-out -- src/module/gcn.py:32-77."""

    def __init__(self, input_dim, hidden_dims, n_layers, dropout=0.5):
        super().__init__()
        _check_uniform(input_dim, hidden_dims, n_layers, "GCN")
        self.dropout_p = dropout
        self.gnn_layers = nn.ModuleList(GCNConv(input_dim) for _ in range(n_layers))
        self.linear_prediction = nn.ModuleList(_head(input_dim, input_dim) for _ in range(n_layers + 1))

    def _flat_params(self):
        cp, hp = [], []
        for conv in self.gnn_layers:
            cp += [conv.ctx_layer.weight, conv.layer_norm.weight, conv.layer_norm.bias]
        for seq in self.linear_prediction:
            hp += _head_params(seq)
        return cp, hp

    def forward(self, x, adj):
        cp, hp = self._flat_params()
        return XF.gnn_layer("GCN", x, adj, cp, hp, self.training, self.dropout_p)

    def forward_regen(self, x, squash=True):
        """forward(x, adj_regen(x)) as ONE autograd node (the generators' layer-to-layer step, ggm.py:221-228)."""
        cp, hp = self._flat_params()
        return XF.gnn_layer("GCN", x, None, cp, hp, self.training, self.dropout_p, regen=squash)


# ---------------------------------------------------------------------------
# GIN  (reference: src/module/gin.py)
# ---------------------------------------------------------------------------
class GINConv(nn.Module):
    """LN(GeLU(Linear(X + ((1+eps) A) @ X))) -- src/module/gin.py:10-34."""

    def __init__(self, input_dim, hidden_dim):
        super().__init__()
        self.eps = nn.Parameter(torch.zeros(1))
        self.linear = _head(input_dim, hidden_dim)

    def forward(self, X, A):
        pre = XF.adj_apply(A, X, 1.0, self.eps, 1.0)
        return _head_forward(self.linear, pre)


class GIN(nn.Module):
    """src/module/gin.py:37-87."""

    def __init__(self, input_dim, hidden_dims, n_layers, dropout=0.5):
        super().__init__()
        _check_uniform(input_dim, hidden_dims, n_layers, "GIN")
        self.dropout_p = dropout
        self.gnn_convs = nn.ModuleList(GINConv(input_dim, input_dim) for _ in range(n_layers))
        self.linear_prediction = nn.ModuleList(_head(input_dim, input_dim) for _ in range(n_layers + 1))

    def _flat_params(self):
        cp, hp = [], []
        for conv in self.gnn_convs:
            cp += [conv.eps] + _head_params(conv.linear)
        for seq in self.linear_prediction:
            hp += _head_params(seq)
        return cp, hp

    def forward(self, X, A):
        cp, hp = self._flat_params()
        return XF.gnn_layer("GIN", X, A, cp, hp, self.training, self.dropout_p)

    def forward_regen(self, X, squash=True):
        cp, hp = self._flat_params()
        return XF.gnn_layer("GIN", X, None, cp, hp, self.training, self.dropout_p, regen=squash)


# ---------------------------------------------------------------------------
# GAT  (reference: src/module/gat.py)
# ---------------------------------------------------------------------------
class GATConv(nn.Module):
    """Dense masked attention -- src/module/gat.py:6-49.  The [B,N,N,2H] concat tensor of the
    reference is never built: a.[h_i || h_j] = a1.h_i + a2.h_j is evaluated per graph in smem."""

    def __init__(self, dim_input, dim_hidden, dropout=0.5, alpha=0.2, concat=True):
        super().__init__()
        self.dropout = dropout
        self.concat = concat
        self.dim_hidden = dim_hidden
        self.alpha = alpha
        self.linear_layer = nn.Linear(dim_input, dim_hidden, bias=False)
        self.attn_layer = nn.Linear(2 * dim_hidden, 1, bias=False)
        self.reset_parameters()
        self.leaky_relu = nn.LeakyReLU(alpha)

    def reset_parameters(self):
        gain = math.sqrt(2.0)  # nn.init.calculate_gain('relu'), src/module/gat.py:20-23
        nn.init.xavier_normal_(self.linear_layer.weight, gain=gain)
        nn.init.xavier_normal_(self.attn_layer.weight, gain=gain)

    def forward(self, x, adj):
        h = XF.linear(x, self.linear_layer.weight)
        return XF.gat_attn(h, self.attn_layer.weight, adj, self.alpha, self.concat)


class GAT(nn.Module):
    """Input dropout + n_head GATConv, concatenated -- src/module/gat.py:52-79."""

    def __init__(self, input_dim, hidden_dim, n_head, dropout=0.5, alpha=0.2, merge='cat'):
        super().__init__()
        self.dropout = dropout
        self.merge = merge
        self.gat_layers = nn.ModuleList(
            GATConv(input_dim, hidden_dim, dropout=dropout, alpha=alpha, concat=True) for _ in range(n_head))

    def forward(self, x, adj):
        x = XF.dropout(x, self.dropout, self.training)
        outs = [att(x, adj) for att in self.gat_layers]
        if self.merge == 'cat':
            return torch.cat(outs, dim=2)
        return torch.mean(torch.stack(outs))  # as the reference (src/module/gat.py:77): a global mean


# ---------------------------------------------------------------------------
# generators  (reference: src/module/graph_generative_modeling.py)
# ---------------------------------------------------------------------------
class _Generator(nn.Module):
    """x <- GNN_l(x, adj); adj <- sigmoid(x x^T / colmax) minus diagonal, per layer."""
    squash = True

    def __init__(self, n_layers, dropout):
        super().__init__()
        self.dropout_p = dropout
        self.n_layers = n_layers

    def forward(self, x, adj):
        for layer in range(self.n_layers):
            if layer > 0 and layer == self.n_layers - 1:
                ddp.notify_layer_boundary(x)   # data-parallel runs: gradients of the last layer can ship early
            gnn = self.gnn_layers[layer]
            if layer > 0 and hasattr(gnn, "forward_regen"):
                # the adjacency between two layers is only ever read by the next layer: regenerate it inside that
                # layer's autograd node (x keeps a single consumer; no gradient-accumulation kernel)
                x = gnn.forward_regen(x, self.squash)
            else:
                if layer > 0:
                    adj = XF.adj_regen(x, self.squash)
                x = gnn(x, adj)
        adj = XF.adj_regen(x, self.squash)
        return x, adj


class GCNGenerator(_Generator):
    """ggm.py:199-233: n_layers x GCN(H, [H,H], n_layers=2)."""

    def __init__(self, hidden_dim, n_layers, dropout=0.5):
        super().__init__(n_layers, dropout)
        self.act = nn.Sigmoid()
        self.gnn_layers = nn.ModuleList(
            GCN(hidden_dim, [hidden_dim, hidden_dim], 2, dropout=dropout) for _ in range(n_layers))


class GINGenerator(_Generator):
    """ggm.py:162-196: n_layers x GIN(H, [H,H], n_layers=1)."""

    def __init__(self, hidden_dim, n_layers, dropout=0.5):
        super().__init__(n_layers, dropout)
        self.act = nn.Sigmoid()
        self.gnn_layers = nn.ModuleList(
            GIN(hidden_dim, [hidden_dim, hidden_dim], 1, dropout=dropout) for _ in range(n_layers))


class GATGenerator(_Generator):
    """ggm.py:236-269: n_layers x GAT(H, H, n_head=2).  As in the reference the 2-head concat
    widens H -> 2H, so only n_layers=1 is shape-valid (layer 2 raises, here as there)."""

    def __init__(self, hidden_dim, n_layers, dropout=0.5):
        super().__init__(n_layers, dropout)
        self.act = nn.Sigmoid()
        self.gnn_layers = nn.ModuleList(GAT(hidden_dim, hidden_dim, n_head=2) for _ in range(n_layers))


class EdgeGenerator(_Generator):
    """ggm.py:100-130: GIN layers, adjacency regenerated WITHOUT the sigmoid; returns adj only."""
    squash = False

    def __init__(self, hidden_dim, n_layers, dropout=0.5):
        super().__init__(n_layers, dropout)
        self.gnn_layers = nn.ModuleList(
            GIN(hidden_dim, [hidden_dim, hidden_dim], 1, dropout=dropout) for _ in range(n_layers))

    def forward(self, x, adj):
        return super().forward(x, adj)[1]


class _PlainStack(nn.Module):
    def __init__(self, n_layers, dropout):
        super().__init__()
        self.dropout_p = dropout
        self.n_layers = n_layers

    def forward(self, x, adj):
        for layer in range(self.n_layers):
            x = self.gnn_layers[layer](x, adj)
        return x


class NodeGenerator(_PlainStack):
    """ggm.py:133-159."""

    def __init__(self, hidden_dim, n_layers, dropout=0.5):
        super().__init__(n_layers, dropout)
        self.gnn_layers = nn.ModuleList(
            GIN(hidden_dim, [hidden_dim, hidden_dim], 1, dropout=dropout) for _ in range(n_layers))


class GinPlainEncoder(_PlainStack):
    """ggm.py:15-40."""

    def __init__(self, hidden_dim, n_layers=2, dropout=0.5):
        super().__init__(n_layers, dropout)
        self.gnn_layers = nn.ModuleList(
            GIN(hidden_dim, [hidden_dim, hidden_dim], 1) for _ in range(n_layers))


class GCNPlainEncoder(_PlainStack):
    """ggm.py:43-68."""

    def __init__(self, hidden_dim, n_layers=2, dropout=0.5):
        super().__init__(n_layers, dropout)
        self.gnn_layers = nn.ModuleList(
            GCN(hidden_dim, [hidden_dim, hidden_dim], 1) for _ in range(n_layers))


class Discriminator(nn.Module):
    """ggm.py:71-82: Linear -> GeLU -> LayerNorm -> Linear on the flattened graph."""

    def __init__(self, hidden_dim):
        super().__init__()
        self.model = nn.Sequential(nn.Linear(hidden_dim, 512), GeLU(), nn.LayerNorm(512), nn.Linear(512, 1))

    def forward(self, x):
        m = self.model
        h = _head_forward(m, x.reshape(x.shape[0], -1))
        return XF.linear(h, m[3].weight, m[3].bias)


class DiscriminatorV2(nn.Module):
    """ggm.py:85-97: Linear -> LeakyReLU(0.2) -> Linear -> LeakyReLU(0.2) -> Linear on the flattened graph.
    (Never built by the shipped trainers; the projections run on the library's GEMM engine, the two
    LeakyReLUs are plain elementwise torch ops.)"""

    def __init__(self, hidden_dim):
        super().__init__()
        self.model = nn.Sequential(nn.Linear(hidden_dim, 512), nn.LeakyReLU(0.2), nn.Linear(512, 256),
                                   nn.LeakyReLU(0.2), nn.Linear(256, 1))

    def forward(self, x):
        m = self.model
        h = m[1](XF.linear(x.reshape(x.shape[0], -1), m[0].weight, m[0].bias))
        h = m[3](XF.linear(h, m[2].weight, m[2].bias))
        return XF.linear(h, m[4].weight, m[4].bias)


class MixGenerator(nn.Module):
    """ggm.py:272-323: VAE-style node sampler (fc1/fc2 -> reparameterise -> decoder to 36 nodes) followed by
    GIN layers; returns (node_feats, rec_loss + kl_div_loss).  Never built by the shipped trainers.  Linear
    layers, LayerNorm and the GIN layers run on the library's kernels; the reparameterisation and the two
    scalar losses are a handful of elementwise torch ops."""

    def __init__(self, hidden_dim, n_layers, dropout=0.5):
        super().__init__()
        self.hidden_dim = hidden_dim
        self.n_layers = n_layers
        self.fc1 = nn.Linear(hidden_dim, hidden_dim)
        self.fc2 = nn.Linear(hidden_dim, hidden_dim)
        self.decoder = nn.Sequential(nn.Linear(hidden_dim, 6 * hidden_dim), nn.LayerNorm(6 * hidden_dim),
                                     nn.ReLU(inplace=True), nn.Linear(6 * hidden_dim, 36 * hidden_dim))
        self.gnn_layers = nn.ModuleList(
            GIN(hidden_dim, [hidden_dim, hidden_dim], 1, dropout=dropout) for _ in range(n_layers))

    def forward(self, x, adj, obj_feats, eps=None):
        node_feats, kl_div_loss = self.generate_node(x, eps)
        rec_loss = torch.nn.functional.binary_cross_entropy_with_logits(node_feats, obj_feats) * 768
        for layer in range(self.n_layers):
            node_feats = self.gnn_layers[layer](node_feats, adj)
        return node_feats, rec_loss + kl_div_loss

    def generate_node(self, x, eps=None):
        mu = XF.linear(x, self.fc1.weight, self.fc1.bias)
        log_var = XF.linear(x, self.fc2.weight, self.fc2.bias)
        z = self.re_parameterize(mu, log_var, eps)
        d = self.decoder
        h = XF.linear(z, d[0].weight, d[0].bias)
        h = torch.relu(XF.layer_norm(h, d[1].weight, d[1].bias, d[1].eps))
        z = XF.linear(h, d[3].weight, d[3].bias).view(-1, 36, self.hidden_dim)
        kl_div_loss = -0.5 * torch.sum(1 + log_var - mu.pow(2) - log_var.exp())
        return z, kl_div_loss

    @staticmethod
    def re_parameterize(mu, log_var, eps=None):
        std = log_var.mul(0.5).exp()
        if eps is None:
            eps = torch.randn_like(std)
        return mu + std * eps


# ---------------------------------------------------------------------------
# VisualFeatEncoder  (reference: src/lxrt/modeling.py:530-556) -- SURVEY 8(f-1): the 2048-d region
# feature projection that produces the graph block's node features
# ---------------------------------------------------------------------------
class VisualFeatEncoder(nn.Module):
    """(LN_1e-12(visn_fc(feats)) + LN_1e-12(box_fc(boxes))) / 2, then dropout.

    ``config`` is the reference's BertConfig (only ``hidden_size`` and ``hidden_dropout_prob`` are read);
    feat_dim / pos_dim are VISUAL_CONFIG.visual_feat_dim / visual_pos_dim (2048 / 4).  Parameter names match
    the reference, so the ``bert.encoder.visn_fc.*`` slice of an LXMERT checkpoint loads unchanged.  The
    2048 -> 768 projection runs on the tcgen05 engine; everything after it (the 4 -> 768 box projection, both
    LayerNorms, the average and the dropout) is ONE row kernel per direction (``xggm_visn_tail_*``)."""

    def __init__(self, config=None, hidden_size=768, hidden_dropout_prob=0.1, feat_dim=2048, pos_dim=4):
        super().__init__()
        if config is not None:
            hidden_size, hidden_dropout_prob = config.hidden_size, config.hidden_dropout_prob
        self.visn_fc = nn.Linear(feat_dim, hidden_size)
        self.visn_layer_norm = nn.LayerNorm(hidden_size, eps=1e-12)
        self.box_fc = nn.Linear(pos_dim, hidden_size)
        self.box_layer_norm = nn.LayerNorm(hidden_size, eps=1e-12)
        self.dropout = nn.Dropout(hidden_dropout_prob)

    def forward(self, visn_input):
        feats, boxes = visn_input
        ln1, ln2 = self.visn_layer_norm, self.box_layer_norm
        if ln1.eps == ln2.eps and XF.visn_tail_supported(self.visn_fc.out_features, self.box_fc.in_features):
            # projection GEMM + ONE fused row kernel (box projection, both LayerNorms, average, dropout)
            z = XF.linear(feats, self.visn_fc.weight, self.visn_fc.bias)
            return XF.visn_tail(z, boxes, self.box_fc.weight, self.box_fc.bias, ln1.weight, ln1.bias, ln2.weight, ln2.bias,
                                self.dropout.p, self.training, ln1.eps)
        x = XF.layer_norm(XF.linear(feats, self.visn_fc.weight, self.visn_fc.bias),
                          self.visn_layer_norm.weight, self.visn_layer_norm.bias, self.visn_layer_norm.eps)
        y = XF.layer_norm(XF.linear(boxes, self.box_fc.weight, self.box_fc.bias),
                          self.box_layer_norm.weight, self.box_layer_norm.bias, self.box_layer_norm.eps)
        return XF.avg2_dropout(x, y, self.dropout.p, self.training)
