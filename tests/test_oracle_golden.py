"""Pins oracle/xggm_oracle.py to the reference: every fixture under tests/golden was
produced by oracle/make_golden.py executing the unmodified reference modules."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_l2, rel_max
from oracle import xggm_oracle as O

TOL = 2e-5  # fp32 oracle vs fp32 reference, different association order only


def _t(a):
    return torch.from_numpy(np.asarray(a))


def _check_param_grads(gold, params, tol):
    seen = 0
    for name, v in params.items():
        g = v.grad if v.grad is not None else torch.zeros_like(v)
        if "g/" + name in gold:
            ref = gold["g/" + name]
            if np.abs(ref).max() == 0:
                assert float(g.abs().max()) == 0.0, name
            else:
                assert rel_l2(g, ref) < tol, name
            seen += 1
        elif "gn/" + name in gold:
            nrm, sm = gold["gn/" + name]
            if nrm == 0:
                assert float(g.abs().max()) == 0.0, name
            else:
                assert abs(float(g.double().norm()) - nrm) / nrm < tol, name
                assert rel_l2(g.reshape(-1)[:16], gold["gh/" + name]) < 50 * tol, name
            seen += 1
    assert seen == sum(1 for k in gold if k.startswith(("g/", "gn/")))


GEN_CASES = [("gcn_h64_train", "GCN"), ("gcn_h64_eval", "GCN"), ("gcn_h768_train", "GCN"),
             ("gin_h64_train", "GIN"), ("gin_h768_train", "GIN"),
             ("gat_h64_train", "GAT"), ("gat_h768_eval", "GAT")]


@pytest.mark.parametrize("name,gnn", GEN_CASES)
def test_generator_matches_reference(name, gnn):
    gold = load_golden(name)
    seed, hidden, B, n_layers, training = [int(v) for v in gold["meta"]]
    p = O.make_params(seed, gnn, hidden, n_layers, 36, heads=False)
    for v in p.values():
        v.requires_grad_(True)
    visn, _, _ = O.make_inputs(seed + 1, B, 36, hidden)
    nh = {"GCN": 3, "GIN": 2, "GAT": 1}[gnn]
    keeps = O.make_keeps(seed + 3, n_layers, nh, (B, 36, hidden)) if training else None
    if gnn == "GAT" and keeps is not None:
        keeps = [k[0] for k in keeps]
    x = visn.clone().requires_grad_(True)
    adj = _t(gold["adj_in"]).clone().requires_grad_(True)
    xo, ao = O.GENERATORS[gnn](x, adj, p, n_layers, keeps, pre="generator.")
    assert rel_l2(xo, gold["x_out"]) < TOL
    assert rel_l2(ao, gold["adj_out"]) < TOL
    # bit-exact structure: zero diagonal
    assert float(torch.diagonal(ao, dim1=1, dim2=2).abs().max()) == 0.0
    ((xo * _t(gold["cx"])).sum() + (ao * _t(gold["ca"])).sum()).backward()
    assert rel_l2(x.grad, gold["gx"]) < 5 * TOL
    ga = adj.grad if adj.grad is not None else torch.zeros_like(adj)
    if np.abs(gold["gadj"]).max() == 0:
        assert float(ga.abs().max()) == 0.0
    else:
        assert rel_l2(ga, gold["gadj"]) < 5 * TOL
    _check_param_grads(gold, p, 10 * TOL)


BRANCH_CASES = [("branch_relation_gcn_h64", "relation", "GCN"),
                ("branch_node_gcn_h64", "node", "GCN"),
                ("branch_node_gcn_h768", "node", "GCN"),
                ("branch_relation_gin_h64", "relation", "GIN"),
                ("branch_relation_gcn_h768", "relation", "GCN")]


@pytest.mark.parametrize("name,which,gnn", BRANCH_CASES)
def test_branch_matches_reference(name, which, gnn):
    gold = load_golden(name)
    seed, hidden, B, n_layers, _ = [int(v) for v in gold["meta"]]
    sigma, A = float(gold["sigma"]), int(gold["num_answers"])
    p = O.make_params(seed, gnn, hidden, n_layers, 36, heads=True)
    for v in p.values():
        v.requires_grad_(True)
    visn, xp, adj_true = O.make_inputs(seed + 1, B, 36, hidden)
    nh = {"GCN": 3, "GIN": 2}[gnn]
    keeps = O.make_keeps(seed + 3, n_layers, nh, (B, 36, hidden))
    xp = xp.clone().requires_grad_(True)
    visn = visn.clone().requires_grad_(True)
    fn = O.relation_branch if which == "relation" else O.node_branch
    x_gen, loss_sm, nodes, adj_g = fn(xp, visn, adj_true, p, sigma, _t(gold["randn"]), keeps, A,
                                      gnn=gnn, n_layers=n_layers)
    assert rel_l2(x_gen, gold["x_gen"]) < TOL
    assert rel_l2(nodes, gold["nodes"]) < TOL
    assert rel_l2(adj_g, gold["adj_gen"]) < TOL
    assert abs(float(loss_sm) - float(gold["loss_sm"])) / abs(float(gold["loss_sm"])) < TOL
    ((x_gen * _t(gold["c"])).sum() + loss_sm).backward()
    assert rel_l2(xp.grad, gold["gxp"]) < 10 * TOL
    gv = visn.grad if visn.grad is not None else torch.zeros_like(visn)
    assert rel_l2(gv, gold["gvisn"]) < 10 * TOL
    _check_param_grads(gold, p, 20 * TOL)


def test_glue_matches_reference():
    g = load_golden("glue")
    a, b, f, h = (_t(g[k]) for k in "abfh")
    an, at = O.edge_noise(a, 0.7, _t(g["rn_a"]))
    fn, ft = O.feat_noise(f, 0.7, _t(g["rn_f"]))
    # noise arithmetic is elementwise in the same order: bit-exact
    assert torch.equal(an, _t(g["edge_noisy"])) and torch.equal(at, _t(g["edge_target"]))
    assert torch.equal(fn, _t(g["feat_noisy"])) and torch.equal(ft, _t(g["feat_target"]))
    assert torch.equal(O.strip_diag(a), _t(g["strip"]))
    assert torch.equal(O.triu_scatter(_t(g["v"]), 36), _t(g["scatter"]))
    assert abs(float(O.score_matching_loss(a, b, 0.7)) - float(g["sm_adj"])) < 1e-6 * abs(float(g["sm_adj"]))
    assert abs(float(O.score_matching_loss(f, h, 0.7)) - float(g["sm_feat"])) < 1e-6 * abs(float(g["sm_feat"]))
    assert abs(float(O.sym_kl_loss(a, b)) - float(g["kl_adj"])) < 1e-5 * abs(float(g["kl_adj"]))
    assert abs(float(O.sym_kl_loss(f, h)) - float(g["kl_feat"])) < 1e-5 * abs(float(g["kl_feat"]))


def test_triu_index_closed_form():
    n = 36
    idx = O.triu_index_map(n)
    for i in range(n):
        for j in range(i + 1, n):
            assert int(idx[i, j]) == i * (2 * n - i - 1) // 2 + (j - i - 1)
    assert int(idx.max()) == n * (n - 1) // 2 - 1


@pytest.mark.parametrize("name", ["visual_feat_train", "visual_feat_eval"])
def test_visual_feat_encoder_matches_reference(name):
    """oracle restatement of lxrt.modeling.VisualFeatEncoder vs the fixture written by the
    unmodified reference class."""
    g = load_golden(name)
    seed, hidden, B, training = [int(v) for v in g["meta"]]
    p = {k: v.requires_grad_(True) for k, v in O.make_visual_params(seed, hidden).items()}
    feats, boxes = O.make_visual_inputs(seed + 1, B)
    f, b = feats.clone().requires_grad_(True), boxes.clone().requires_grad_(True)
    keep = _t(g["keep"]).bool() if training else None
    out = O.visual_feat_encoder(f, b, p, keep=keep, drop_p=float(g["drop_p"]))
    assert rel_l2(out, g["out"]) < TOL
    (out * _t(g["c"])).sum().backward()
    assert abs(float(f.grad.double().norm()) - float(g["gfeats_norm"][0])) < 10 * TOL * float(g["gfeats_norm"][0])
    assert rel_l2(f.grad.reshape(-1)[:64], g["gfeats_head"]) < 20 * TOL
    assert rel_l2(b.grad, g["gboxes"]) < 20 * TOL
    for k in ("box_fc.weight", "box_fc.bias", "visn_layer_norm.weight", "box_layer_norm.bias"):
        assert rel_l2(p[k].grad, g["g/" + k]) < 20 * TOL, k


def test_optimizer_tail_oracle_matches_reference_fixture():
    """BertAdam + clip_grad_norm_(5.) + BCEWithLogits*A restated in oracle/ vs the reference classes executed by
    oracle/make_golden.py (src/lxrt/optimization.py:116-203, src/vqa/vqacpv2.py:110,173-177)."""
    g = load_golden("optimizer")
    steps = int(g["seed"][1])
    n = sum(1 for k in g if k.startswith("p0/"))
    lr, t_total, warmup = float(g["lr"][0]), int(g["t_total"][0]), float(g["warmup"][0])
    wd, max_norm = float(g["weight_decay"][0]), float(g["max_norm"][0])
    p = [_t(g[f"p0/{i}"]) for i in range(n)]
    m = [torch.zeros_like(t) for t in p]
    v = [torch.zeros_like(t) for t in p]
    clipped = 0
    for s in range(steps):
        grads = [_t(g[f"g{s}/{i}"]) for i in range(n)]
        norm, coef = O.clip_coef(grads, max_norm)
        clipped += coef < 1.0
        assert abs(norm - float(g[f"norm{s}"][0])) <= 1e-5 * norm
        lr_s = O.scheduled_lr(lr, s, t_total, warmup)
        assert abs(lr_s - float(g[f"lr{s}"][0])) <= 1e-12
        for i in range(n):
            p[i], m[i], v[i] = O.bertadam_step(p[i], grads[i] * coef, m[i], v[i], lr_s, weight_decay=wd)
            assert rel_l2(p[i], g[f"p{s + 1}/{i}"]) < 1e-6 and rel_l2(m[i], g[f"m{s + 1}/{i}"]) < 1e-6
            assert rel_l2(v[i], g[f"v{s + 1}/{i}"]) < 1e-6
    assert 0 < clipped < steps   # the fixture exercises both the clipped and the unclipped regime
    x = _t(g["bce/logit"]).requires_grad_(True)
    loss = O.bce_with_logits(x, _t(g["bce/target"]), scale=x.shape[1])
    loss.backward()
    assert abs(float(loss) - float(g["bce/loss"][0])) <= 1e-5 * abs(float(g["bce/loss"][0]))
    assert rel_l2(x.grad, g["bce/glogit"]) < 1e-5


# ---------------------------------------------------------------------------- SURVEY 8 a-18
API_CASES = ["edge_generator", "node_generator", "gin_plain_encoder", "gcn_plain_encoder", "discriminator",
             "discriminator_v2", "gcn_conv_dropout", "gat_mean"]


def api_surface_oracle(tag, gold, dtype=torch.float32):
    """Run the oracle's restatement of one API-surface class on the fixture's inputs, weights and masks.
    Returns (y, x, adj, params) with gradients populated for cotangent gold[tag/c]."""
    seed, hidden, B, n_layers = [int(v) for v in gold["meta"]]
    p = {k[len(tag) + 4:]: _t(v).to(dtype).requires_grad_(True) for k, v in gold.items() if k.startswith(tag + "/sd/")}
    x = _t(gold["x_in"]).to(dtype).requires_grad_(True)
    adj = _t(gold["adj_in"]).to(dtype).requires_grad_(True)
    masks = []
    while f"{tag}/keep{len(masks)}" in gold:
        masks.append(_t(gold[f"{tag}/keep{len(masks)}"]))
    keeps = [masks[2 * l:2 * l + 2] for l in range(n_layers)]
    if tag == "edge_generator":
        y = O.edge_generator(x, adj, p, n_layers, keeps)
    elif tag in ("node_generator", "gin_plain_encoder"):
        y = O.gnn_stack(x, adj, p, "GIN", n_layers, keeps)
    elif tag == "gcn_plain_encoder":
        y = O.gnn_stack(x, adj, p, "GCN", n_layers, keeps)
    elif tag == "discriminator":
        y = O.discriminator(x[:, :2], p)
    elif tag == "discriminator_v2":
        y = O.discriminator_v2(x[:, :2], p)
    elif tag == "gcn_conv_dropout":
        y = O.gcn_conv_dropout(x, adj, p, "", masks[0], 0.25)
    elif tag == "gat_mean":
        y = O.gat_mean(x, adj, p, "", 2, masks[0]).reshape(1)
    else:
        raise KeyError(tag)
    (y * _t(gold[tag + "/c"]).to(dtype)).sum().backward()
    return y, x, adj, p


@pytest.mark.parametrize("tag", API_CASES)
def test_api_surface_matches_reference(tag):
    gold = load_golden("api_surface")
    y, x, adj, p = api_surface_oracle(tag, gold)
    assert rel_l2(y, gold[tag + "/y"]) < TOL
    assert rel_l2(x.grad, gold[tag + "/gx"]) < 5 * TOL
    if np.abs(gold[tag + "/gadj"]).max() > 0:
        assert rel_l2(adj.grad, gold[tag + "/gadj"]) < 5 * TOL
    else:
        assert adj.grad is None or float(adj.grad.abs().max()) == 0.0
    n = 0
    for k, v in p.items():
        ref = gold[f"{tag}/g/{k}"]
        g = v.grad if v.grad is not None else torch.zeros_like(v)
        if np.abs(ref).max() == 0:
            assert float(g.abs().max()) == 0.0, k
        else:
            assert rel_l2(g, ref) < 10 * TOL, k
        n += 1
    assert n == sum(1 for k in gold if k.startswith(tag + "/g/"))


@pytest.mark.parametrize("branch", ["node", "relation"])
def test_reference_step_runner(branch):
    """oracle/ref_step.py (the baselines' step: real reference modules from oracle/_ref when vendored, else the
    port) runs a full training step on CPU and moves the parameters."""
    from oracle.ref_step import ReferenceStep, reference_available
    kinds = [True, False] if reference_available() else [False]
    for prefer in kinds:
        s = ReferenceStep("cpu", 2, branch, hidden=64, prefer_reference=prefer, lr=1e-3)
        assert s.kind == ("reference" if prefer else "port")
        if s.kind == "reference":
            before = [p.detach().clone() for p in s.params]
        l1 = float(s.step())
        l2 = float(s.step())
        assert l1 == l1 and l2 == l2 and abs(l1) < 1e6
        if s.kind == "reference":
            moved = sum(float((p.detach() - b).abs().sum()) for p, b in zip(s.params, before))
            assert moved > 0
