"""CPU oracle for the X-GGM graph block -- TEST INFRASTRUCTURE, NOT PRODUCT.

This file restates, as plain functions over a flat ``{name: tensor}`` parameter
dictionary, the algorithm of the reference's graph-generative hot path so that the
CUDA path in ``xggm_b200/`` can be checked against it on any machine (the
reference tree itself does not travel to the GPU box).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it; nothing under ``xggm_b200/`` does.

Parity status: PINNED.  The reference ships no tests or golden vectors
(SURVEY.md section 4), so the pin is the reference itself executed in the build
container: ``oracle/make_golden.py`` imports the unmodified reference modules from
``/root/reference/src`` and writes ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function below against those files.

All arithmetic is torch CPU in the dtype of the inputs (fp32 or fp64).  Stochastic
pieces (dropout keep-masks, Gaussian noise) are *injected* so they can be shared
with the CUDA path bit for bit.

Reference citations are relative to ``/root/reference``.
"""
import math

import torch

LN_EPS = 1e-5  # nn.LayerNorm default, src/module/gcn.py:14,47
NEG_FILL = -9e15  # src/module/gat.py:40


# --------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------
def gelu_erf(t):
    """Exact-erf GeLU.  src/lxrt/modeling.py:116-124."""
    return t * 0.5 * (1.0 + torch.erf(t / math.sqrt(2.0)))


def row_norm(t, gamma, beta, eps=LN_EPS):
    """LayerNorm over the last axis (biased variance): nn.LayerNorm, as the reference uses."""
    return torch.nn.functional.layer_norm(t, (t.shape[-1],), gamma, beta, eps)


def affine(t, weight, bias=None):
    """y = t W^T (+ b): nn.Linear."""
    return torch.nn.functional.linear(t, weight, bias)


def keep_scale(t, keep, p):
    """Inverted dropout with an explicit keep-mask (1 = keep).  F.dropout semantics
    (src/module/gcn.py:72-76): kept values are scaled by 1/(1-p)."""
    if keep is None:
        return t
    return t * keep.to(t.dtype) / (1.0 - p)


def strip_diag(a):
    """a.triu(1) + a.tril(-1).  src/vqa/vqacpv2.py:188, ggm.py:228."""
    n = a.shape[-1]
    off = 1.0 - torch.eye(n, dtype=a.dtype)
    return a * off


# --------------------------------------------------------------------------
# VisualFeatEncoder  (src/lxrt/modeling.py:530-556) -- SURVEY 8(f-1), the producer of the block's input
# --------------------------------------------------------------------------
BERT_LN_EPS = 1e-12  # BertLayerNorm(config.hidden_size, eps=1e-12), modeling.py:538,542


def visual_feat_encoder(feats, boxes, p, pre="", keep=None, drop_p=0.1):
    """(LN(visn_fc(feats)) + LN(box_fc(boxes))) / 2, then dropout.  modeling.py:546-556."""
    x = row_norm(affine(feats, p[pre + "visn_fc.weight"], p[pre + "visn_fc.bias"]),
                 p[pre + "visn_layer_norm.weight"], p[pre + "visn_layer_norm.bias"], BERT_LN_EPS)
    y = row_norm(affine(boxes, p[pre + "box_fc.weight"], p[pre + "box_fc.bias"]),
                 p[pre + "box_layer_norm.weight"], p[pre + "box_layer_norm.bias"], BERT_LN_EPS)
    return keep_scale((x + y) / 2, keep, drop_p)


def make_visual_params(seed, hidden=768, feat_dim=2048, pos_dim=4, dtype=torch.float32):
    """Seeded parameters of VisualFeatEncoder (nn.Linear-style uniform init; LN affine perturbed so that
    gamma/beta gradients are exercised)."""
    g = torch.Generator().manual_seed(seed)

    def uni(shape, fan_in):
        b = 1.0 / math.sqrt(fan_in)
        return ((torch.rand(shape, generator=g) * 2 - 1) * b).to(dtype)

    return {
        "visn_fc.weight": uni((hidden, feat_dim), feat_dim), "visn_fc.bias": uni((hidden,), feat_dim),
        "visn_layer_norm.weight": (1 + 0.1 * torch.randn(hidden, generator=g)).to(dtype),
        "visn_layer_norm.bias": (0.1 * torch.randn(hidden, generator=g)).to(dtype),
        "box_fc.weight": uni((hidden, pos_dim), pos_dim), "box_fc.bias": uni((hidden,), pos_dim),
        "box_layer_norm.weight": (1 + 0.1 * torch.randn(hidden, generator=g)).to(dtype),
        "box_layer_norm.bias": (0.1 * torch.randn(hidden, generator=g)).to(dtype),
    }


def make_visual_inputs(seed, B, n_obj=36, feat_dim=2048, dtype=torch.float32):
    """feats = relu(N(0,1)) (post-ReLU pool5 look-alike), boxes = sorted U(0,1) corners."""
    g = torch.Generator().manual_seed(seed)
    feats = torch.relu(torch.randn(B, n_obj, feat_dim, generator=g)).to(dtype)
    xy = torch.rand(B, n_obj, 2, 2, generator=g).sort(dim=-1)[0]          # [..., axis, (lo, hi)]
    boxes = torch.stack([xy[..., 0, 0], xy[..., 1, 0], xy[..., 0, 1], xy[..., 1, 1]], dim=-1).to(dtype)  # x1,y1,x2,y2
    return feats, boxes


# --------------------------------------------------------------------------
# GCN  (src/module/gcn.py)
# --------------------------------------------------------------------------
def gcn_conv(x, adj, p, pre):
    """GCNConv.forward, src/module/gcn.py:22-29 (dropout p=0 is the identity)."""
    agg = torch.bmm(adj, x)
    u = x + affine(agg, p[pre + "ctx_layer.weight"])
    return row_norm(u, p[pre + "layer_norm.weight"], p[pre + "layer_norm.bias"])


def jk_head(h, p, pre):
    """One ``linear_prediction[j]`` = Linear -> GeLU -> LayerNorm.
    src/module/gcn.py:44-62."""
    z = affine(h, p[pre + "0.weight"], p[pre + "0.bias"])
    return row_norm(gelu_erf(z), p[pre + "2.weight"], p[pre + "2.bias"])


def gcn(x, adj, p, pre="", n_convs=2, keeps=None, drop_p=0.5):
    """GCN.forward, src/module/gcn.py:64-77.  ``keeps`` is a list of n_convs+1
    keep-masks (or None for eval mode)."""
    hs = [x]
    for k in range(n_convs):
        hs.append(gcn_conv(hs[-1], adj, p, f"{pre}gnn_layers.{k}."))
    out = 0.0
    for j, h in enumerate(hs):
        y = jk_head(h, p, f"{pre}linear_prediction.{j}.")
        out = out + keep_scale(y, None if keeps is None else keeps[j], drop_p)
    return out


def adj_regen(x, squash=True):
    """Adjacency regeneration, ggm.py:225-228 (squash=False: EdgeGenerator
    ggm.py:124-126).  Note the indexing: the column max m[b, i] = max_k S[b, k, i]
    divides *row* i."""
    s = torch.bmm(x, x.transpose(1, 2))
    m = s.max(dim=1)[0].unsqueeze(-1)
    a = s / m
    if squash:
        a = torch.sigmoid(a)
    return strip_diag(a)


def gcn_generator(x, adj, p, n_layers=2, keeps=None, drop_p=0.5, pre=""):
    """GCNGenerator.forward, ggm.py:214-233.  keeps[l][j]."""
    for l in range(n_layers):
        x = gcn(x, adj, p, f"{pre}gnn_layers.{l}.", 2,
                None if keeps is None else keeps[l], drop_p)
        adj = adj_regen(x)
    return x, adj


# --------------------------------------------------------------------------
# GIN  (src/module/gin.py)
# --------------------------------------------------------------------------
def gin_conv(x, adj, p, pre):
    """GINConv.forward, src/module/gin.py:21-34: X + ((1+eps) A) @ X, then
    Linear -> GeLU -> LN."""
    x = x + torch.bmm((1 + p[pre + "eps"]) * adj, x)
    z = affine(x, p[pre + "linear.0.weight"], p[pre + "linear.0.bias"])
    return row_norm(gelu_erf(z), p[pre + "linear.2.weight"], p[pre + "linear.2.bias"])


def gin(x, adj, p, pre="", n_convs=1, keeps=None, drop_p=0.5):
    """GIN.forward, src/module/gin.py:68-87."""
    hs = [x]
    for k in range(n_convs):
        hs.append(gin_conv(hs[-1], adj, p, f"{pre}gnn_convs.{k}."))
    out = 0.0
    for j, h in enumerate(hs):
        y = jk_head(h, p, f"{pre}linear_prediction.{j}.")
        out = out + keep_scale(y, None if keeps is None else keeps[j], drop_p)
    return out


def gin_generator(x, adj, p, n_layers=2, keeps=None, drop_p=0.5, pre=""):
    """GINGenerator.forward, ggm.py:177-196."""
    for l in range(n_layers):
        x = gin(x, adj, p, f"{pre}gnn_layers.{l}.", 1,
                None if keeps is None else keeps[l], drop_p)
        adj = adj_regen(x)
    return x, adj


# --------------------------------------------------------------------------
# GAT  (src/module/gat.py)
# --------------------------------------------------------------------------
def gat_conv(x, adj, p, pre, alpha=0.2):
    """GATConv.forward, src/module/gat.py:25-49.  The concat attention
    a.[h_i || h_j] is evaluated as (a1.h_i) + (a2.h_j), which is the same sum in a
    different association order (differences are fp rounding only)."""
    h = affine(x, p[pre + "linear_layer.weight"])
    a = p[pre + "attn_layer.weight"].reshape(-1)
    d = h.shape[-1]
    s_self = h @ a[:d]
    s_nbr = h @ a[d:]
    e = torch.nn.functional.leaky_relu(s_self.unsqueeze(2) + s_nbr.unsqueeze(1), alpha)
    e = e.masked_fill(adj == 0, NEG_FILL)
    att = torch.softmax(e, dim=-1)
    return torch.nn.functional.elu(torch.bmm(att, h))


def gat(x, adj, p, pre="", n_head=2, keep=None, drop_p=0.5):
    """GAT.forward (merge='cat'), src/module/gat.py:72-79."""
    x = keep_scale(x, keep, drop_p)
    return torch.cat([gat_conv(x, adj, p, f"{pre}gat_layers.{h}.")
                      for h in range(n_head)], dim=2)


def gat_generator(x, adj, p, n_layers=1, keeps=None, drop_p=0.5, pre=""):
    """GATGenerator.forward, ggm.py:250-269 (only n_layers=1 is shape-valid in
    the reference: the 2-head concat widens 768 -> 1536)."""
    for l in range(n_layers):
        x = gat(x, adj, p, f"{pre}gnn_layers.{l}.", 2,
                None if keeps is None else keeps[l], drop_p)
        adj = adj_regen(x)
    return x, adj


# --------------------------------------------------------------------------
# API-surface compositions no trainer builds: ggm.py:15-159
# --------------------------------------------------------------------------
def gcn_conv_dropout(x, adj, p, pre, keep, drop_p):
    """GCNConv.forward with dropout > 0 on the projected aggregate, src/module/gcn.py:28."""
    agg = affine(torch.bmm(adj, x), p[pre + "ctx_layer.weight"])
    u = x + keep_scale(agg, keep, drop_p)
    return row_norm(u, p[pre + "layer_norm.weight"], p[pre + "layer_norm.bias"])


def gnn_stack(x, adj, p, kind, n_layers, keeps=None, drop_p=0.5, pre=""):
    """GinPlainEncoder / GCNPlainEncoder / NodeGenerator.forward (ggm.py:27-40, 55-68, 146-159): n_layers x
    GIN|GCN(n_layers=1) -- one conv + two read-out heads each -- applied in sequence with the SAME adjacency.
    keeps[l][j]."""
    layer = gin if kind == "GIN" else gcn
    for l in range(n_layers):
        x = layer(x, adj, p, f"{pre}gnn_layers.{l}.", 1, None if keeps is None else keeps[l], drop_p)
    return x


def edge_generator(x, adj, p, n_layers, keeps=None, drop_p=0.5, pre=""):
    """EdgeGenerator.forward, ggm.py:115-130: GIN(n_layers=1) layers, adjacency regenerated WITHOUT the
    sigmoid; returns the adjacency only."""
    for l in range(n_layers):
        x = gin(x, adj, p, f"{pre}gnn_layers.{l}.", 1, None if keeps is None else keeps[l], drop_p)
        adj = adj_regen(x, squash=False)
    return adj


def discriminator(x, p, pre="model."):
    """Discriminator.forward, ggm.py:71-82: Linear -> GeLU -> LayerNorm -> Linear on the flattened input."""
    h = x.reshape(x.shape[0], -1)
    h = row_norm(gelu_erf(affine(h, p[pre + "0.weight"], p[pre + "0.bias"])), p[pre + "2.weight"], p[pre + "2.bias"])
    return affine(h, p[pre + "3.weight"], p[pre + "3.bias"])


def discriminator_v2(x, p, pre="model."):
    """DiscriminatorV2.forward, ggm.py:85-97."""
    lrelu = torch.nn.functional.leaky_relu
    h = x.reshape(x.shape[0], -1)
    h = lrelu(affine(h, p[pre + "0.weight"], p[pre + "0.bias"]), 0.2)
    h = lrelu(affine(h, p[pre + "2.weight"], p[pre + "2.bias"]), 0.2)
    return affine(h, p[pre + "4.weight"], p[pre + "4.bias"])


def gat_mean(x, adj, p, pre="", n_head=2, keep=None, drop_p=0.5):
    """GAT.forward with merge != 'cat', src/module/gat.py:76-77: torch.mean over the STACK of head outputs --
    a global scalar mean (all heads, graphs, nodes and features), as the reference computes it."""
    x = keep_scale(x, keep, drop_p)
    return torch.mean(torch.stack([gat_conv(x, adj, p, f"{pre}gat_layers.{h}.") for h in range(n_head)]))


# --------------------------------------------------------------------------
# noise + losses + trainer glue
# --------------------------------------------------------------------------
def edge_noise(adj, sigma, randn):
    """add_edge_noise_v2, src/module/graph_utils.py:162-168; ``randn`` stands in
    for torch.randn_like(adj)."""
    n = randn.triu(diagonal=1) * sigma
    n = n + n.transpose(-1, -2)
    return adj + n, -n / (sigma ** 2)


def feat_noise(f, sigma, randn):
    """add_feature_noise_v2, src/module/graph_utils.py:144-149."""
    n = randn * sigma
    return f + n, -n / (sigma ** 2)


def score_matching_loss(score, target, sigma):
    """loss_func, src/vqa/vqacpv2.py:48-51."""
    cur = 0.5 * sigma ** 2 * ((score - target) ** 2).sum(dim=[-1, -2]).mean()
    return cur / (score.shape[-1] * score.shape[-2])


def sym_kl_loss(x, y):
    """compute_kl_loss, src/vqa/vqacpv2.py:54-61:
    mean over all elements of py (log py - log px) + px (log px - log py)."""
    lpx = torch.log_softmax(x, dim=-1)
    lpy = torch.log_softmax(y, dim=-1)
    px, py = lpx.exp(), lpy.exp()
    return (py * (lpy - lpx) + px * (lpx - lpy)).mean()


def triu_index_map(n):
    """Row-major (i<j) enumeration used by the boolean-mask assignment at
    src/vqa/vqacpv2.py:195-198: k = i(2n-i-1)/2 + (j-i-1)."""
    idx = torch.full((n, n), -1, dtype=torch.long)
    k = 0
    for i in range(n):
        for j in range(i + 1, n):
            idx[i, j] = k
            idx[j, i] = k
            k += 1
    return idx


def triu_scatter(v, n):
    """v[B, n(n-1)/2] -> symmetric [B,n,n] with zero diagonal.
    src/vqa/vqacpv2.py:195-199."""
    idx = triu_index_map(n)
    out = torch.zeros(v.shape[0], n, n, dtype=v.dtype)
    off = idx >= 0
    out[:, off] = v[:, idx[off]]
    return out


def adj_encoder(xp, p, pre="encoder_adj."):
    """encoder_adj = Linear(768,630)+Sigmoid, src/vqa/vqacpv2_model.py:91-94."""
    return torch.sigmoid(affine(xp, p[pre + "0.weight"], p[pre + "0.bias"]))


def node_init(xp, n, p, pre="node_fc."):
    """node_fc(x.unsqueeze(1).repeat(1,n,1)), src/vqa/vqacpv2.py:228-229 and
    src/vqa/vqacpv2_model.py:95-99 (Linear -> GeLU -> LN on n identical rows)."""
    return jk_head(xp.unsqueeze(1).repeat(1, n, 1), p, pre)


def fusion_readout(xp, nodes, p, pre="fusion_fc."):
    """fusion_fc(cat[x, tanh(mean_n nodes)]), src/vqa/vqacpv2.py:216-218 and
    src/vqa/vqacpv2_model.py:101-105."""
    cat = torch.cat([xp, torch.tanh(nodes.mean(1))], dim=-1)
    return jk_head(cat, p, pre)


GENERATORS = {"GCN": gcn_generator, "GIN": gin_generator, "GAT": gat_generator}


def relation_branch(xp, visn, adj_true, p, sigma, randn, keeps, num_answers,
                    gnn="GCN", n_layers=2, kl_weight=8.0):
    """Relation-generation GGM step up to (x_gen, loss_sm).
    src/vqa/vqacpv2.py:194-218 (kl_weight 8) / src/gqa/gqa_ood.py:178-202 (12).
    ``adj_true`` is the raw obj36_adj matrix (the diagonal is stripped here)."""
    n = adj_true.shape[-1]
    adj_t = strip_diag(adj_true)
    adj0 = triu_scatter(adj_encoder(xp, p), n)
    adj_n, tgt = edge_noise(adj0, sigma, randn)
    nodes, adj_g = GENERATORS[gnn](visn, adj_n, p, n_layers, keeps, pre="generator.")
    l_sm = score_matching_loss(adj_g, tgt, sigma)
    l_kl = sym_kl_loss(adj_t, adj_g) * num_answers
    loss_sm = kl_weight * l_kl + l_sm
    x_gen = fusion_readout(xp, nodes, p)
    return x_gen, loss_sm, nodes, adj_g


def node_branch(xp, visn, adj_true, p, sigma, randn, keeps, num_answers,
                gnn="GCN", n_layers=2):
    """Node (representation) generation GGM step up to (x_gen, loss_sm).
    src/vqa/vqacpv2.py:226-247 / src/gqa/gqa_ood.py:236-252."""
    n = adj_true.shape[-1]
    adj_t = strip_diag(adj_true)
    nodes0 = node_init(xp, n, p)
    nodes_n, tgt = feat_noise(nodes0, sigma, randn)
    nodes, adj_g = GENERATORS[gnn](nodes_n, adj_t, p, n_layers, keeps, pre="generator.")
    l_kl = sym_kl_loss(nodes, visn) * num_answers
    l_sm = score_matching_loss(nodes, tgt, sigma)
    loss_sm = 0.15 * l_kl + 6 * l_sm
    x_gen = fusion_readout(xp, nodes, p)
    return x_gen, loss_sm, nodes, adj_g


# --------------------------------------------------------------------------
# deterministic synthetic inputs / parameters
# --------------------------------------------------------------------------
def param_shapes(gnn="GCN", hidden=768, n_layers=2, n_nodes=36, heads=True):
    """Names and shapes of the block's parameters, in reference state_dict order
    (probed from the reference modules)."""
    H = hidden
    out = []
    for l in range(n_layers):
        g = f"generator.gnn_layers.{l}."
        if gnn == "GCN":
            for k in range(2):
                out += [(g + f"gnn_layers.{k}.ctx_layer.weight", (H, H)),
                        (g + f"gnn_layers.{k}.layer_norm.weight", (H,)),
                        (g + f"gnn_layers.{k}.layer_norm.bias", (H,))]
            nh = 3
        elif gnn == "GIN":
            out += [(g + "gnn_convs.0.eps", (1,)),
                    (g + "gnn_convs.0.linear.0.weight", (H, H)),
                    (g + "gnn_convs.0.linear.0.bias", (H,)),
                    (g + "gnn_convs.0.linear.2.weight", (H,)),
                    (g + "gnn_convs.0.linear.2.bias", (H,))]
            nh = 2
        elif gnn == "GAT":
            for h in range(2):
                out += [(g + f"gat_layers.{h}.linear_layer.weight", (H, H)),
                        (g + f"gat_layers.{h}.attn_layer.weight", (1, 2 * H))]
            nh = 0
        else:
            raise KeyError(gnn)
        for j in range(nh):
            out += [(g + f"linear_prediction.{j}.0.weight", (H, H)),
                    (g + f"linear_prediction.{j}.0.bias", (H,)),
                    (g + f"linear_prediction.{j}.2.weight", (H,)),
                    (g + f"linear_prediction.{j}.2.bias", (H,))]
    if heads:
        E = n_nodes * (n_nodes - 1) // 2
        out += [("encoder_adj.0.weight", (E, H)), ("encoder_adj.0.bias", (E,)),
                ("node_fc.0.weight", (H, H)), ("node_fc.0.bias", (H,)),
                ("node_fc.2.weight", (H,)), ("node_fc.2.bias", (H,)),
                ("fusion_fc.0.weight", (H, 2 * H)), ("fusion_fc.0.bias", (H,)),
                ("fusion_fc.2.weight", (H,)), ("fusion_fc.2.bias", (H,))]
    return out


def make_params(seed, gnn="GCN", hidden=768, n_layers=2, n_nodes=36, heads=True,
                dtype=torch.float32):
    """Seeded parameters (NOT the reference's init order -- a self-contained
    recipe so fixtures only need to store the seed): matrices U(-1,1)/sqrt(fan_in),
    vectors 1 + 0.1 N(0,1) for LN gains, 0.1 N(0,1) otherwise, eps 0.1."""
    g = torch.Generator().manual_seed(seed)
    p = {}
    for name, shape in param_shapes(gnn, hidden, n_layers, n_nodes, heads):
        if len(shape) == 2:
            t = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) / math.sqrt(shape[1])
        elif name.endswith("eps"):
            t = torch.full(shape, 0.1, dtype=torch.float64)
        elif name.endswith("2.weight") or name.endswith("layer_norm.weight"):
            t = 1 + 0.1 * torch.randn(shape, generator=g, dtype=torch.float64)
        else:
            t = 0.1 * torch.randn(shape, generator=g, dtype=torch.float64)
        p[name] = t.to(dtype)
    return p


def make_inputs(seed, B, n_nodes=36, hidden=768, dtype=torch.float32):
    """Synthetic block inputs in the shapes of SURVEY.md section 8d: visn =
    layer_norm(N(0,1)); pooled = tanh(N(0,1)); adj_true = obj36_adj look-alike
    (symmetrised U(-0.2,1), max-normalised, non-zero diagonal)."""
    g = torch.Generator().manual_seed(seed)
    v = torch.randn(B, n_nodes, hidden, generator=g, dtype=torch.float64)
    v = (v - v.mean(-1, keepdim=True)) / v.std(-1, unbiased=False, keepdim=True)
    xp = torch.tanh(torch.randn(B, hidden, generator=g, dtype=torch.float64))
    c = torch.rand(B, n_nodes, n_nodes, generator=g, dtype=torch.float64) * 1.2 - 0.2
    c = c + c.transpose(1, 2)
    c = c / c.amax(dim=(1, 2), keepdim=True)
    return v.to(dtype), xp.to(dtype), c.to(dtype)


def make_keeps(seed, n_layers, n_heads, shape, p=0.5):
    """Bernoulli(1-p) keep-masks as uint8, keeps[l][j]."""
    g = torch.Generator().manual_seed(seed)
    return [[(torch.rand(shape, generator=g) >= p).to(torch.uint8) for _ in range(n_heads)]
            for _ in range(n_layers)]


# ---------------------------------------------------------------------------
# training-step tail
# ---------------------------------------------------------------------------
def warmup_constant(x, warmup=0.002):
    """src/lxrt/optimization.py:34-40."""
    return x / warmup if x < warmup else 1.0


def warmup_linear(x, warmup=0.002):
    """src/lxrt/optimization.py:43-49."""
    return x / warmup if x < warmup else max((x - 1.0) / (warmup - 1.0), 0)


def warmup_cosine(x, warmup=0.002):
    """src/lxrt/optimization.py:28-31 (the reference calls torch.cos on a Python float there; math.cos is the
    intended arithmetic)."""
    return x / warmup if x < warmup else 0.5 * (1.0 + math.cos(math.pi * x))


SCHEDULES = {"warmup_cosine": warmup_cosine, "warmup_constant": warmup_constant, "warmup_linear": warmup_linear}


def scheduled_lr(lr, step, t_total, warmup, schedule="warmup_linear"):
    """BertAdam's per-step learning rate, src/lxrt/optimization.py:176-191 (`step` = steps already taken)."""
    if t_total == -1:
        return lr
    return lr * SCHEDULES[schedule](step / t_total, warmup)


def clip_coef(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_(params, max_norm) as the trainers call it (src/vqa/vqacpv2.py:175):
    total 2-norm over ALL gradients, coefficient max_norm / (norm + 1e-6) clamped to 1.  Returns (norm, coef)."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads))
    return float(total), min(1.0, max_norm / (float(total) + 1e-6))


def bertadam_step(p, g, m, v, lr, b1=0.9, b2=0.999, e=1e-6, weight_decay=0.01):
    """One BertAdam update of one tensor, src/lxrt/optimization.py:156-193 (no bias correction; `lr` already
    scheduled; `g` already clipped).  Returns (p, m, v) without modifying the inputs."""
    m = m * b1 + (1 - b1) * g
    v = v * b2 + (1 - b2) * g * g
    update = m / (v.sqrt() + e)
    if weight_decay > 0.0:
        update = update + weight_decay * p
    return p - lr * update, m, v


def bce_with_logits(x, t, scale=1.0):
    """nn.BCEWithLogitsLoss()(x, t) * scale (src/vqa/vqacpv2.py:110,173): mean over all elements of
    max(x,0) - x t + log(1 + exp(-|x|))."""
    return scale * (x.clamp(min=0) - x * t + torch.log1p(torch.exp(-x.abs()))).mean()
