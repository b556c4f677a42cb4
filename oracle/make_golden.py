"""Generate tests/golden/*.npz by executing the UNMODIFIED reference.

Runs only in the build container (needs /root/reference); the fixtures it writes
are committed so the GPU box never needs the reference tree.

    python oracle/make_golden.py

What is executed from the reference:
  * module.graph_generative_modeling.{GCN,GIN,GAT}Generator, module.gcn/gin/gat,
    module.graph_utils.add_{edge,feature}_noise_v2 -- imported as they are
    (boto3/botocore, which lxrt.file_utils imports for downloads, are stubbed in
    sys.modules; no reference file is edited or copied);
  * loss_func / compute_kl_loss -- the FunctionDef nodes are cut out of
    src/vqa/vqacpv2.py with ``ast`` and exec'd (the module itself cannot be
    imported: param.py parses argv, h5py/tensorboardX are absent).
The trainer lines between them (src/vqa/vqacpv2.py:187-218,226-247) are glue with
no callable entry point and are followed line by line in ``_branch`` below.

Parameters/inputs/masks come from oracle.xggm_oracle.make_* (seeded), so the
fixtures store seeds + outputs only.
"""
import ast
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference/src"
sys.path.insert(0, ROOT)

from oracle import xggm_oracle as O  # noqa: E402


def _import_reference():
    for m in ("boto3", "botocore", "botocore.exceptions"):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules["botocore.exceptions"].ClientError = Exception
    sys.path.insert(0, REF)
    import module.graph_generative_modeling as ggm
    import module.graph_utils as gu
    src = open(os.path.join(REF, "vqa", "vqacpv2.py")).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "F": torch.nn.functional}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("loss_func", "compute_kl_loss"):
            exec(compile(ast.Module([node], []), "vqacpv2.py", "exec"), ns)
    return ggm, gu, ns["loss_func"], ns["compute_kl_loss"]


class _MaskFeeder:
    """Replaces torch.nn.functional.dropout while a reference module runs so the
    keep-masks are the injected ones, consumed in call order."""

    def __init__(self, masks, p):
        self.masks, self.p, self.i = list(masks), p, 0

    def __call__(self, x, p=0.5, training=True, inplace=False):
        if not training or p == 0.0:
            return x
        assert abs(p - self.p) < 1e-12
        k = self.masks[self.i].to(x.dtype)
        self.i += 1
        return x * k / (1.0 - p)


def _run_with_masks(fn, masks, p=0.5):
    F = torch.nn.functional
    orig = F.dropout
    feeder = _MaskFeeder(masks, p)
    F.dropout = feeder
    try:
        out = fn()
    finally:
        F.dropout = orig
    assert feeder.i == len(masks), (feeder.i, len(masks))
    return out


def _load(module, params, prefix):
    sd = {k[len(prefix):]: v.clone() for k, v in params.items() if k.startswith(prefix)}
    missing = module.load_state_dict(sd, strict=True)
    return missing


def _np(t):
    return t.detach().cpu().numpy()


def _grad_summary(named, full):
    out = {}
    for k, v in named:
        g = v.grad if v.grad is not None else torch.zeros_like(v)
        if full:
            out["g/" + k] = _np(g)
        else:
            flat = g.reshape(-1)
            out["gn/" + k] = np.array([float(flat.double().norm()), float(flat.double().sum())])
            out["gh/" + k] = _np(flat[:16])
    return out


def generator_case(ggm, gnn, hidden, B, n_layers, seed, training=True, dtype=torch.float32):
    N = 36
    p = O.make_params(seed, gnn, hidden, n_layers, N, heads=False, dtype=dtype)
    visn, _, adj_true = O.make_inputs(seed + 1, B, N, hidden, dtype)
    g = torch.Generator().manual_seed(seed + 2)
    adj = O.strip_diag(adj_true) + 0.3 * torch.randn(B, N, N, generator=g).to(dtype)
    if gnn == "GAT":  # exercise the adj == 0 mask (src/module/gat.py:40)
        adj = adj * (torch.rand(B, N, N, generator=g) > 0.3).to(dtype)
    cls = {"GCN": ggm.GCNGenerator, "GIN": ggm.GINGenerator, "GAT": ggm.GATGenerator}[gnn]
    mod = cls(hidden_dim=hidden, n_layers=n_layers).to(dtype)
    _load(mod, p, "generator.")
    mod.train(training)
    nh = {"GCN": 3, "GIN": 2, "GAT": 1}[gnn]
    keeps = O.make_keeps(seed + 3, n_layers, nh, (B, N, hidden)) if training else None
    x = visn.clone().requires_grad_(True)
    a = adj.clone().requires_grad_(True)
    flat_masks = [m for layer in keeps for m in layer] if training else []
    xo, ao = _run_with_masks(lambda: mod(x, a), flat_masks)
    cx = torch.randn(xo.shape, generator=g).to(dtype)
    ca = torch.randn(ao.shape, generator=g).to(dtype)
    loss = (xo * cx).sum() + (ao * ca).sum()
    loss.backward()
    out = dict(meta=np.array([seed, hidden, B, n_layers, int(training)]),
               adj_in=_np(adj), cx=_np(cx), ca=_np(ca),
               x_out=_np(xo), adj_out=_np(ao), gx=_np(x.grad),
               gadj=_np(a.grad if a.grad is not None else torch.zeros_like(a)))
    out.update(_grad_summary([("generator." + k, v) for k, v in mod.named_parameters()],
                             full=hidden <= 64))
    return out


def _branch(ggm, gu, loss_func, compute_kl_loss, which, gnn, hidden, B, seed, sigma=1.0,
            num_answers=2274, n_layers=2, dtype=torch.float32):
    """Follows src/vqa/vqacpv2.py:187-218 (relation) and :226-247 (node) using the
    reference's own modules and loss functions."""
    import torch.nn as nn
    from lxrt.modeling import GeLU
    N = 36
    E = N * (N - 1) // 2
    p = O.make_params(seed, gnn, hidden, n_layers, N, heads=True, dtype=dtype)
    visn, xp, adj_true = O.make_inputs(seed + 1, B, N, hidden, dtype)
    cls = {"GCN": ggm.GCNGenerator, "GIN": ggm.GINGenerator}[gnn]
    generator = cls(hidden_dim=hidden, n_layers=n_layers).to(dtype)
    # heads as built at src/vqa/vqacpv2_model.py:91-105
    encoder_adj = nn.Sequential(nn.Linear(hidden, E), nn.Sigmoid()).to(dtype)
    node_fc = nn.Sequential(nn.Linear(hidden, hidden), GeLU(), nn.LayerNorm(hidden)).to(dtype)
    fusion_fc = nn.Sequential(nn.Linear(hidden * 2, hidden), GeLU(), nn.LayerNorm(hidden)).to(dtype)
    _load(generator, p, "generator.")
    _load(encoder_adj, p, "encoder_adj.")
    _load(node_fc, p, "node_fc.")
    _load(fusion_fc, p, "fusion_fc.")
    for m in (generator, encoder_adj, node_fc, fusion_fc):
        m.train()
    nh = {"GCN": 3, "GIN": 2}[gnn]
    keeps = O.make_keeps(seed + 3, n_layers, nh, (B, N, hidden))
    flat_masks = [m for layer in keeps for m in layer]
    g = torch.Generator().manual_seed(seed + 4)
    x = xp.clone().requires_grad_(True)
    feat = visn.clone().requires_grad_(True)
    adj_t = adj_true.triu(1) + adj_true.tril(-1)          # vqacpv2.py:188
    orig_randn_like = torch.randn_like
    if which == "relation":
        randn = torch.randn(B, N, N, generator=g).to(dtype)
        adj_noise = torch.zeros_like(adj_t)                 # :196
        adj_temp = torch.ones_like(adj_t).triu(1)           # :197
        adj_noise[adj_temp == 1] = encoder_adj(x).view(-1)  # :198
        adj_noise = adj_noise + adj_noise.transpose(1, 2)   # :199
        torch.randn_like = lambda t: randn
        try:
            adj_noise, grad_log_noise = gu.add_edge_noise_v2(adj_noise, sigma=sigma)  # :201
        finally:
            torch.randn_like = orig_randn_like
        node_feats, adj_gen = _run_with_masks(lambda: generator(feat, adj_noise), flat_masks)
        loss_grad = loss_func(adj_gen, grad_log_noise, sigma=sigma)    # :208
        d_loss = compute_kl_loss(adj_t, adj_gen) * num_answers         # :210
        loss_sm = 8 * d_loss + loss_grad                               # :212
    else:
        randn = torch.randn(B, N, hidden, generator=g).to(dtype)
        node_feats = x.unsqueeze(1).repeat(1, 36, 1)                   # :228
        node_feats = node_fc(node_feats)                               # :229
        torch.randn_like = lambda t: randn
        try:
            node_feats, feat_grad = gu.add_feature_noise_v2(node_feats, sigma=sigma)  # :230
        finally:
            torch.randn_like = orig_randn_like
        node_feats, adj_gen = _run_with_masks(lambda: generator(node_feats, adj_t), flat_masks)
        d_loss = compute_kl_loss(node_feats, feat) * num_answers       # :237
        loss_grad = loss_func(node_feats, feat_grad, sigma=sigma)      # :239
        loss_sm = 0.15 * d_loss + 6 * loss_grad                        # :241
    x_gen = fusion_fc(torch.cat([x, torch.tanh(node_feats.mean(1))], dim=-1))  # :216 / :245
    c = torch.randn(x_gen.shape, generator=g).to(dtype)
    loss = (x_gen * c).sum() + loss_sm
    loss.backward()
    named = [("generator." + k, v) for k, v in generator.named_parameters()]
    named += [("encoder_adj." + k, v) for k, v in encoder_adj.named_parameters()]
    named += [("node_fc." + k, v) for k, v in node_fc.named_parameters()]
    named += [("fusion_fc." + k, v) for k, v in fusion_fc.named_parameters()]
    out = dict(meta=np.array([seed, hidden, B, n_layers, 1]), sigma=np.array(sigma),
               num_answers=np.array(num_answers), randn=_np(randn), c=_np(c),
               x_gen=_np(x_gen), loss_sm=_np(loss_sm), d_loss=_np(d_loss),
               loss_grad=_np(loss_grad), nodes=_np(node_feats), adj_gen=_np(adj_gen),
               gxp=_np(x.grad), gvisn=_np(feat.grad if feat.grad is not None else torch.zeros_like(feat)))
    out.update(_grad_summary(named, full=hidden <= 64))
    return out


def api_surface_case(ggm, seed=501, hidden=64, B=3, n_layers=2):
    """the classes of ggm.py:15-159,272-323 no trainer builds, plus GCNConv(dropout>0) (gcn.py:28) and
    GAT(merge != 'cat') (gat.py:76-77).  Each reference module is built under a seed, its own state_dict is
    stored, and it runs forward + backward on shared inputs with injected dropout masks."""
    import module.gcn as gcn_mod
    import module.gat as gat_mod
    N = 36
    g = torch.Generator().manual_seed(seed)
    visn, _, adj_true = O.make_inputs(seed + 1, B, N, hidden)
    adj = O.strip_diag(adj_true) + 0.3 * torch.randn(B, N, N, generator=g)
    out = {"x_in": _np(visn), "adj_in": _np(adj), "meta": np.array([seed, hidden, B, n_layers])}

    def run(tag, mod, fn, n_masks, mask_shape, p=0.5):
        torch.manual_seed(seed + len(tag))
        for k_, v_ in mod.named_parameters():      # perturb the default init so LN / eps gradients are exercised
            if v_.dim() == 1:
                v_.data.add_(0.1 * torch.randn(v_.shape, generator=g))
        mod.train()
        masks = [(torch.rand(mask_shape, generator=g) >= p).to(torch.uint8) for _ in range(n_masks)]
        x = visn.clone().requires_grad_(True)
        a = adj.clone().requires_grad_(True)
        y = _run_with_masks(lambda: fn(mod, x, a), masks, p)
        c = torch.randn(y.shape, generator=g)
        (y * c).sum().backward()
        for k_, v_ in mod.state_dict().items():
            out[f"{tag}/sd/{k_}"] = _np(v_)
        for i, m in enumerate(masks):
            out[f"{tag}/keep{i}"] = _np(m)
        out[f"{tag}/c"] = _np(c)
        out[f"{tag}/y"] = _np(y)
        out[f"{tag}/gx"] = _np(x.grad if x.grad is not None else torch.zeros_like(x))
        out[f"{tag}/gadj"] = _np(a.grad if a.grad is not None else torch.zeros_like(a))
        for k_, v_ in mod.named_parameters():
            out[f"{tag}/g/{k_}"] = _np(v_.grad if v_.grad is not None else torch.zeros_like(v_))

    shape = (B, N, hidden)
    run("edge_generator", ggm.EdgeGenerator(hidden, n_layers), lambda m, x, a: m(x, a), 2 * n_layers, shape)
    run("node_generator", ggm.NodeGenerator(hidden, n_layers), lambda m, x, a: m(x, a), 2 * n_layers, shape)
    run("gin_plain_encoder", ggm.GinPlainEncoder(hidden, n_layers), lambda m, x, a: m(x, a), 2 * n_layers, shape)
    run("gcn_plain_encoder", ggm.GCNPlainEncoder(hidden, n_layers), lambda m, x, a: m(x, a), 2 * n_layers, shape)
    # (the discriminators flatten their input; two node rows = 2*hidden features keep the fixture small)
    run("discriminator", ggm.Discriminator(2 * hidden), lambda m, x, a: m(x[:, :2]), 0, shape)
    run("discriminator_v2", ggm.DiscriminatorV2(2 * hidden), lambda m, x, a: m(x[:, :2]), 0, shape)
    run("gcn_conv_dropout", gcn_mod.GCNConv(hidden, dropout=0.25), lambda m, x, a: m(x, a), 1, shape, p=0.25)
    run("gat_mean", gat_mod.GAT(hidden, hidden, n_head=2, merge="mean"), lambda m, x, a: m(x, a).reshape(1), 1, shape)
    return out


def glue_case(gu, loss_func, compute_kl_loss, seed=7):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(3, 36, 36, generator=g)
    b = torch.randn(3, 36, 36, generator=g)
    f = torch.randn(3, 36, 96, generator=g)
    h = torch.randn(3, 36, 96, generator=g)
    rn_a = torch.randn(3, 36, 36, generator=g)
    rn_f = torch.randn(3, 36, 96, generator=g)
    orig = torch.randn_like
    try:
        torch.randn_like = lambda t: rn_a
        an, at = gu.add_edge_noise_v2(a, sigma=0.7)
        torch.randn_like = lambda t: rn_f
        fn, ft = gu.add_feature_noise_v2(f, sigma=0.7)
    finally:
        torch.randn_like = orig
    # the boolean-mask scatter of src/vqa/vqacpv2.py:195-199
    v = torch.rand(3, 630, generator=g)
    sc = torch.zeros(3, 36, 36)
    sc[torch.ones(3, 36, 36).triu(1) == 1] = v.view(-1)
    sc = sc + sc.transpose(1, 2)
    return dict(a=_np(a), b=_np(b), f=_np(f), h=_np(h), rn_a=_np(rn_a), rn_f=_np(rn_f),
                edge_noisy=_np(an), edge_target=_np(at), feat_noisy=_np(fn), feat_target=_np(ft),
                sm_adj=_np(loss_func(a, b, sigma=0.7)), sm_feat=_np(loss_func(f, h, sigma=0.7)),
                kl_adj=_np(compute_kl_loss(a, b)), kl_feat=_np(compute_kl_loss(f, h)),
                strip=_np(a.triu(1) + a.tril(-1)), v=_np(v), scatter=_np(sc))


def visual_feat_case(seed=301, B=3, hidden=768, training=True):
    """lxrt.modeling.VisualFeatEncoder (src/lxrt/modeling.py:530-556), imported unmodified."""
    from lxrt.modeling import BertConfig, VisualFeatEncoder
    cfg = BertConfig(vocab_size_or_config_json_file=30522, hidden_size=hidden)
    mod = VisualFeatEncoder(cfg)
    p = O.make_visual_params(seed, hidden)
    mod.load_state_dict({k: v.clone() for k, v in p.items()}, strict=True)
    mod.train(training)
    feats, boxes = O.make_visual_inputs(seed + 1, B)
    g = torch.Generator().manual_seed(seed + 2)
    keep = (torch.rand(B, 36, hidden, generator=g) >= cfg.hidden_dropout_prob)
    f = feats.clone().requires_grad_(True)
    b = boxes.clone().requires_grad_(True)
    out = _run_with_masks(lambda: mod((f, b)), [keep] if training else [], p=cfg.hidden_dropout_prob)
    c = torch.randn(out.shape, generator=g)
    (out * c).sum().backward()
    res = dict(meta=np.array([seed, hidden, B, int(training)]), drop_p=np.array(cfg.hidden_dropout_prob),
               keep=_np(keep.to(torch.uint8)), c=_np(c), out=_np(out),
               gfeats_norm=np.array([float(f.grad.double().norm())]), gfeats_head=_np(f.grad.reshape(-1)[:64]),
               gboxes=_np(b.grad))
    res.update(_grad_summary(list(mod.named_parameters()), full=False))
    for k in ("box_fc.weight", "box_fc.bias", "visn_layer_norm.weight", "box_layer_norm.bias"):
        res["g/" + k] = _np(dict(mod.named_parameters())[k].grad)
    return res


def _npc(t):
    return t.detach().cpu().numpy().copy()   # a snapshot: the optimiser keeps updating these tensors in place


def optimizer_case(seed=401, steps=4):
    """The reference BertAdam (src/lxrt/optimization.py, imported as is) driven the way the trainers drive it:
    clip_grad_norm_(params, 5.) then optim.step() (src/vqa/vqacpv2.py:175-177), warmup_linear schedule, plus
    nn.BCEWithLogitsLoss()(logit, target) * target.size(1) and its gradient (src/vqa/vqacpv2.py:110,173)."""
    sys.path.insert(0, REF)
    from lxrt.optimization import BertAdam
    g = torch.Generator().manual_seed(seed)
    shapes = [(48, 40), (40,), (7, 33), (1,)]
    params = [torch.nn.Parameter(torch.randn(s, generator=g) * 0.3) for s in shapes]
    res = {"seed": np.array([seed, steps]), "lr": np.array([4e-3]), "t_total": np.array([10]),
           "warmup": np.array([0.25]), "weight_decay": np.array([0.01]), "max_norm": np.array([5.0])}
    for i, p in enumerate(params):
        res[f"p0/{i}"] = _npc(p)
    opt = BertAdam(params, lr=4e-3, warmup=0.25, t_total=10)
    for s in range(steps):
        scale = [0.05, 3.0, 0.5, 40.0][s % 4]          # steps with and without active clipping
        for i, p in enumerate(params):
            p.grad = torch.randn(p.shape, generator=g) * scale
            res[f"g{s}/{i}"] = _npc(p.grad)
        norm = torch.nn.utils.clip_grad_norm_(params, 5.0)
        res[f"norm{s}"] = np.array([float(norm)])
        res[f"lr{s}"] = np.array(opt.get_lr()[:1] if s else [0.0])   # get_lr() is [0] before the first step
        opt.step()
        for i, p in enumerate(params):
            res[f"p{s + 1}/{i}"] = _npc(p)
            res[f"m{s + 1}/{i}"] = _npc(opt.state[p]["next_m"])
            res[f"v{s + 1}/{i}"] = _npc(opt.state[p]["next_v"])
    logit = (torch.randn(5, 2274, generator=g) * 3).requires_grad_(True)
    target = (torch.rand(5, 2274, generator=g) < 0.002).float() * torch.tensor([0.3, 0.6, 0.9, 1.0])[
        torch.randint(0, 4, (5, 2274), generator=g)]
    loss = torch.nn.BCEWithLogitsLoss()(logit, target) * target.size(1)
    loss.backward()
    res.update({"bce/logit": _npc(logit), "bce/target": _npc(target), "bce/loss": np.array([float(loss)]),
                "bce/glogit": _npc(logit.grad)})
    return res


def main():
    torch.set_num_threads(1)  # fixed summation order for the fixtures
    ggm, gu, loss_func, compute_kl_loss = _import_reference()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    cases = {
        "gcn_h64_train": lambda: generator_case(ggm, "GCN", 64, 3, 2, 101),
        "gcn_h64_eval": lambda: generator_case(ggm, "GCN", 64, 3, 2, 102, training=False),
        "gcn_h768_train": lambda: generator_case(ggm, "GCN", 768, 2, 2, 103),
        "gin_h64_train": lambda: generator_case(ggm, "GIN", 64, 3, 2, 104),
        "gin_h768_train": lambda: generator_case(ggm, "GIN", 768, 2, 2, 105),
        "gat_h64_train": lambda: generator_case(ggm, "GAT", 64, 3, 1, 106),
        "gat_h768_eval": lambda: generator_case(ggm, "GAT", 768, 2, 1, 107, training=False),
        "branch_relation_gcn_h64": lambda: _branch(ggm, gu, loss_func, compute_kl_loss, "relation", "GCN", 64, 3, 201),
        "branch_node_gcn_h64": lambda: _branch(ggm, gu, loss_func, compute_kl_loss, "node", "GCN", 64, 3, 202),
        "branch_node_gcn_h768": lambda: _branch(ggm, gu, loss_func, compute_kl_loss, "node", "GCN", 768, 2, 203),
        "branch_relation_gin_h64": lambda: _branch(ggm, gu, loss_func, compute_kl_loss, "relation", "GIN", 64, 3, 204),
        "branch_relation_gcn_h768": lambda: _branch(ggm, gu, loss_func, compute_kl_loss, "relation", "GCN", 768, 2, 205),
        "api_surface": lambda: api_surface_case(ggm),
        "glue": lambda: glue_case(gu, loss_func, compute_kl_loss),
        "visual_feat_train": lambda: visual_feat_case(301, 3, 768, True),
        "visual_feat_eval": lambda: visual_feat_case(302, 2, 768, False),
        "optimizer": lambda: optimizer_case(),
    }
    only = [a for a in sys.argv[1:] if not a.startswith("-")]   # regenerate just the named cases
    for name, fn in cases.items():
        if only and name not in only:
            continue
        data = fn()
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **data)
        print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
