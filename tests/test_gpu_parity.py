"""GPU parity: the CUDA path (through the C ABI / nn.Module boundary) against the CPU oracle
and the committed golden fixtures.  fp32 tolerance 1e-4 relative (BASELINE.json north_star);
masks, index maps, arg-max indices and noise arithmetic are bit-exact."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_l2, rel_max
from oracle import xggm_oracle as O

pytestmark = pytest.mark.gpu
ROOT_DIR = __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))

TOL = 1e-4


def dev():
    return torch.device("cuda:0")


def _t(a):
    return torch.from_numpy(np.asarray(a))


def _close(a, b, tol=TOL, name=""):
    b = torch.as_tensor(b)
    if float(b.abs().max()) == 0.0:
        assert float(a.detach().abs().max().cpu()) == 0.0, name
        return
    e2, em = rel_l2(a.detach().cpu(), b), rel_max(a.detach().cpu(), b)
    assert e2 < tol and em < 5 * tol, f"{name}: rel_l2={e2:.3e} rel_max={em:.3e}"


def _load_params(module, params, prefix=""):
    sd = {k[len(prefix):]: v for k, v in params.items() if k.startswith(prefix)}
    module.load_state_dict(sd, strict=True)
    return module


# --------------------------------------------------------------------------- primitives
@pytest.mark.parametrize("M,N,K", [(72, 768, 768), (3, 630, 768), (130, 64, 96), (1, 768, 1536), (257, 10, 7)])
def test_linear_fwd_bwd(M, N, K):
    import xggm_b200.functional as XF
    g = torch.Generator().manual_seed(M * 7 + N)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    r = torch.randn(M, N, generator=g)
    c = torch.randn(M, N, generator=g)
    ad, wd, bd, rd = (t.clone().to(dev()).requires_grad_(True) for t in (a, w, b, r))
    out = XF.linear(ad, wd, bd, rd)
    (out * c.to(dev())).sum().backward()
    a64, w64, b64, r64 = (t.double().requires_grad_(True) for t in (a, w, b, r))
    ref = a64 @ w64.T + b64 + r64
    (ref * c.double()).sum().backward()
    _close(out, ref, name="out")
    _close(ad.grad, a64.grad, name="ga")
    _close(wd.grad, w64.grad, name="gw")
    _close(bd.grad, b64.grad, name="gb")
    _close(rd.grad, r64.grad, name="gr")


@pytest.mark.parametrize("B,N,H", [(3, 36, 768), (2, 5, 64), (1, 64, 96), (2, 100, 40)])
def test_adj_apply_and_regen(B, N, H):
    import xggm_b200.functional as XF
    g = torch.Generator().manual_seed(B * 100 + N)
    x = torch.randn(B, N, H, generator=g)
    adj = torch.randn(B, N, N, generator=g)
    eps = torch.tensor([0.25])
    cx = torch.randn(B, N, H, generator=g)
    ca = torch.randn(B, N, N, generator=g)
    xd, ad, ed = (t.clone().to(dev()).requires_grad_(True) for t in (x, adj, eps))
    out = XF.adj_apply(ad, xd, 1.0, ed, 1.0)
    (out * cx.to(dev())).sum().backward()
    x64, a64, e64 = (t.double().requires_grad_(True) for t in (x, adj, eps))
    ref = x64 + torch.bmm((1 + e64) * a64, x64)
    (ref * cx.double()).sum().backward()
    _close(out, ref, name="adj_apply")
    _close(xd.grad, x64.grad, name="gx")
    _close(ad.grad, a64.grad, name="gadj")
    _close(ed.grad, e64.grad, name="geps")
    # regeneration, both variants
    for squash in (True, False):
        xd2 = x.clone().to(dev()).requires_grad_(True)
        adj_g, amax = XF.adj_regen(xd2, squash, return_argmax=True)
        (adj_g * ca.to(dev())).sum().backward()
        x2 = x.double().requires_grad_(True)
        ref_g = O.adj_regen(x2, squash)
        (ref_g * ca.double()).sum().backward()
        _close(adj_g, ref_g, name="regen")
        _close(xd2.grad, x2.grad, name="regen gx")
        assert float(torch.diagonal(adj_g, dim1=1, dim2=2).abs().max()) == 0.0  # bit-exact mask
        # bit-exact arg-max row per column against fp32 torch on the kernel's own S ordering
        s32 = torch.bmm(x, x.transpose(1, 2))
        ref_arg = s32.max(dim=1)[1]
        s_top2 = s32.topk(2, dim=1)[0]
        decisive = (s_top2[:, 0] - s_top2[:, 1]) > 1e-4 * s_top2[:, 0].abs()
        assert torch.equal(amax.cpu().long()[decisive], ref_arg[decisive])


def test_regen_symmetry_and_ties():
    """S is bitwise symmetric (same k-order for (i,j) and (j,i)); ties pick the first row."""
    import xggm_b200.functional as XF
    x = torch.randn(4, 36, 768, generator=torch.Generator().manual_seed(5))
    x[0, 7] = x[0, 3]  # duplicate node: S[3,c] == S[7,c] for every c -> tie between rows 3 and 7
    x[0, 3] *= 4.0
    x[0, 7] *= 4.0
    adj, amax = XF.adj_regen(x.to(dev()), False, return_argmax=True)
    # squash=False: adj[i,j] = S[i,j]/m_i off-diagonal. Row-scale back and compare transposes.
    assert torch.equal(amax[0, 3].cpu(), torch.tensor(3, dtype=torch.int32))
    assert int(amax[0, 7]) == 3  # first index on the exact tie, as torch.max
    a = adj.cpu()
    s = torch.bmm(x, x.transpose(1, 2))
    m = s.max(dim=1)[0]
    i, j = 2, 9
    assert abs(float(a[1, i, j] * m[1, i]) - float(a[1, j, i] * m[1, j])) < 1e-3 * abs(float(s[1, i, j])) + 1e-5


@pytest.mark.parametrize("M,H", [(72, 768), (5, 64), (33, 1536), (1, 32)])
def test_rowops(M, H):
    import xggm_b200.functional as XF
    g = torch.Generator().manual_seed(M + H)
    z = torch.randn(M, H, generator=g) * 2
    gam = 1 + 0.2 * torch.randn(H, generator=g)
    bet = 0.2 * torch.randn(H, generator=g)
    keep = (torch.rand(M, H, generator=g) > 0.5).to(torch.uint8)
    c = torch.randn(M, H, generator=g)
    # LayerNorm
    zd, gd, bd = (t.clone().to(dev()).requires_grad_(True) for t in (z, gam, bet))
    out = XF.layer_norm(zd, gd, bd)
    (out * c.to(dev())).sum().backward()
    z64, g64, b64 = (t.double().requires_grad_(True) for t in (z, gam, bet))
    ref = O.row_norm(z64, g64, b64)
    (ref * c.double()).sum().backward()
    for a, b, n in ((out, ref, "ln"), (zd.grad, z64.grad, "gz"), (gd.grad, g64.grad, "ggamma"), (bd.grad, b64.grad, "gbeta")):
        _close(a, b, name=n)
    # dropout(LN(GeLU(z)))
    zd, gd, bd = (t.clone().to(dev()).requires_grad_(True) for t in (z, gam, bet))
    out = XF.gelu_ln_drop(zd, gd, bd, keep.to(dev()), 0.5)
    (out * c.to(dev())).sum().backward()
    z64, g64, b64 = (t.double().requires_grad_(True) for t in (z, gam, bet))
    ref = O.keep_scale(O.row_norm(O.gelu_erf(z64), g64, b64), keep, 0.5)
    (ref * c.double()).sum().backward()
    for a, b, n in ((out, ref, "head"), (zd.grad, z64.grad, "gz"), (gd.grad, g64.grad, "ggamma"), (bd.grad, b64.grad, "gbeta")):
        _close(a, b, name=n)
    assert torch.equal((out == 0).cpu() | (keep == 1), torch.ones(M, H, dtype=torch.bool))  # dropped entries are exactly 0
    # GeLU alone
    zd = z.clone().to(dev()).requires_grad_(True)
    y = XF.gelu(zd)
    (y * c.to(dev())).sum().backward()
    z64 = z.double().requires_grad_(True)
    r = O.gelu_erf(z64)
    (r * c.double()).sum().backward()
    _close(y, r, name="gelu")
    _close(zd.grad, z64.grad, name="gelu grad")


def test_glue_bit_exact_against_reference_fixture():
    import xggm_b200 as X
    g = load_golden("glue")
    d = dev()
    a, f = _t(g["a"]).to(d), _t(g["f"]).to(d)
    an, at = X.add_edge_noise(a, sigma=0.7, randn=_t(g["rn_a"]).to(d))
    fn, ft = X.add_feature_noise(f, sigma=0.7, randn=_t(g["rn_f"]).to(d))
    assert torch.equal(an.cpu(), _t(g["edge_noisy"])) and torch.equal(at.cpu(), _t(g["edge_target"]))
    assert torch.equal(fn.cpu(), _t(g["feat_noisy"])) and torch.equal(ft.cpu(), _t(g["feat_target"]))
    assert torch.equal(X.glue.strip_diag(a).cpu(), _t(g["strip"]))
    assert torch.equal(X.glue.triu_scatter(_t(g["v"]).to(d), 36).cpu(), _t(g["scatter"]))
    b, h = _t(g["b"]).to(d), _t(g["h"]).to(d)
    for got, key in ((X.loss_func(a, b, sigma=0.7), "sm_adj"), (X.loss_func(f, h, sigma=0.7), "sm_feat"),
                     (X.compute_kl_loss(a, b), "kl_adj"), (X.compute_kl_loss(f, h), "kl_feat")):
        assert abs(float(got) - float(g[key])) < 1e-5 * abs(float(g[key])), key


def test_glue_gradients():
    import xggm_b200 as X
    g = torch.Generator().manual_seed(11)
    d = dev()
    for shape in ((3, 36, 36), (2, 36, 768), (4, 7, 50)):
        x, y = torch.randn(shape, generator=g), torch.randn(shape, generator=g)
        xd, yd = x.clone().to(d).requires_grad_(True), y.clone().to(d).requires_grad_(True)
        (3.0 * X.compute_kl_loss(xd, yd) + 2.0 * X.loss_func(xd, yd, sigma=0.5)).backward()
        x64, y64 = x.double().requires_grad_(True), y.double().requires_grad_(True)
        (3.0 * O.sym_kl_loss(x64, y64) + 2.0 * O.score_matching_loss(x64, y64, 0.5)).backward()
        _close(xd.grad, x64.grad, name="gx")
        _close(yd.grad, y64.grad, name="gy")
    # scatter / readout / broadcast-noise backward
    v = torch.rand(3, 630, generator=g)
    c = torch.randn(3, 36, 36, generator=g)
    vd = v.clone().to(d).requires_grad_(True)
    (X.glue.triu_scatter(vd, 36) * c.to(d)).sum().backward()
    v64 = v.double().requires_grad_(True)
    (O.triu_scatter(v64, 36) * c.double()).sum().backward()
    assert torch.equal(vd.grad.cpu(), v64.grad.float())  # a two-term sum: exact in fp32
    xp, nodes = torch.randn(3, 64, generator=g), torch.randn(3, 36, 64, generator=g)
    c2 = torch.randn(3, 128, generator=g)
    xpd, nd = xp.clone().to(d).requires_grad_(True), nodes.clone().to(d).requires_grad_(True)
    (X.glue.fuse_readout(xpd, nd) * c2.to(d)).sum().backward()
    xp64, n64 = xp.double().requires_grad_(True), nodes.double().requires_grad_(True)
    (torch.cat([xp64, torch.tanh(n64.mean(1))], -1) * c2.double()).sum().backward()
    _close(xpd.grad, xp64.grad, name="gxp")
    _close(nd.grad, n64.grad, name="gnodes")
    f = torch.randn(3, 64, generator=g)
    rn = torch.randn(3, 36, 64, generator=g)
    fd = f.clone().to(d).requires_grad_(True)
    noisy, tgt = X.add_feature_noise(fd, sigma=1.0, randn=rn.to(d))
    (noisy * nodes.to(d)).sum().backward()
    _close(fd.grad, nodes.sum(1), name="broadcast grad")
    assert torch.equal(noisy.cpu(), f.unsqueeze(1) + rn)


@pytest.mark.parametrize("B,N,H", [(3, 36, 768), (2, 5, 64), (4, 7, 50), (1, 36, 128)])
def test_node_tail_matches_oracle_pieces(B, N, H):
    """xggm_node_tail_* (fused KL + score matching + read-out) vs the fp64 oracle of the three pieces
    (src/vqa/vqacpv2.py:236-246), both the vectorised (H % 128 == 0) and the generic kernels."""
    import xggm_b200.functional as XF
    g = torch.Generator().manual_seed(B * 100 + H)
    nodes, feat, tgt = (torch.randn(B, N, H, generator=g) for _ in range(3))
    xp = torch.randn(B, H, generator=g)
    c = torch.randn(B, 2 * H, generator=g)
    sigma, kl_w, sm_w = 0.7, 0.15 * 2274, 6.0
    d = dev()
    nd, fd, xd = (t.clone().to(d).requires_grad_(True) for t in (nodes, feat, xp))
    loss, cat = XF.node_tail(nd, fd, tgt.to(d), xd, sigma, kl_w, sm_w)
    (1.1 * loss + (cat * c.to(d)).sum()).backward()
    n64, f64, x64 = (t.double().requires_grad_(True) for t in (nodes, feat, xp))
    ref_loss = kl_w * O.sym_kl_loss(n64, f64) + sm_w * O.score_matching_loss(n64, tgt.double(), sigma)
    ref_cat = torch.cat([x64, torch.tanh(n64.mean(1))], -1)
    (1.1 * ref_loss + (ref_cat * c.double()).sum()).backward()
    assert abs(float(loss) - float(ref_loss)) <= TOL * abs(float(ref_loss))
    _close(cat, ref_cat, name="cat")
    _close(nd.grad, n64.grad, name="gnodes")
    _close(fd.grad, f64.grad, name="gfeat")
    _close(xd.grad, x64.grad, name="gxp")


# --------------------------------------------------------------------------- generators
GEN_CASES = [("gcn_h64_train", "GCN"), ("gcn_h64_eval", "GCN"), ("gcn_h768_train", "GCN"),
             ("gin_h64_train", "GIN"), ("gin_h768_train", "GIN"),
             ("gat_h64_train", "GAT"), ("gat_h768_eval", "GAT")]


# GIN's eps gradient is a single scalar <gpre, adj @ h> summed with heavy cancellation over
# B*N*H terms: torch's own fp32 result differs from fp64 by 1.2e-4 on the h768 fixture.  The
# split-bf16 tensor-core engine (~3x the rounding error of native fp32 per product) is therefore
# checked at 2e-3 on that one ill-conditioned scalar; the exact SIMT engine keeps the strict bound.
SCALAR_TOL_TC = 2e-3


def _param_grad_check(gold, named, tol, scalar_tol=None):
    base_tol = tol
    for name, v in named:
        tol = scalar_tol if (scalar_tol is not None and v.numel() == 1) else base_tol
        g = v.grad if v.grad is not None else torch.zeros_like(v)
        if "g/" + name in gold:
            _close(g, gold["g/" + name], tol, name)
        elif "gn/" + name in gold:
            nrm = float(gold["gn/" + name][0])
            if nrm == 0:
                assert float(g.abs().max()) == 0.0, name
            else:
                assert abs(float(g.double().norm()) - nrm) / nrm < tol, name
                _close(g.reshape(-1)[:16], gold["gh/" + name], 20 * tol, name)
        else:
            raise AssertionError("fixture has no gradient for " + name)


@pytest.fixture(params=["fp32", "fp32_simt"])
def engine(request):
    """Run under the tcgen05 split-bf16 engine (default) and the exact SIMT engine."""
    import xggm_b200 as X
    X.set_precision(request.param)
    yield request.param
    X.set_precision("fp32")


@pytest.mark.parametrize("name,gnn", GEN_CASES)
def test_generator_matches_reference_fixture(name, gnn, engine):
    import xggm_b200 as X
    from xggm_b200.functional import inject_keep_masks
    gold = load_golden(name)
    seed, hidden, B, n_layers, training = [int(v) for v in gold["meta"]]
    p = O.make_params(seed, gnn, hidden, n_layers, 36, heads=False)
    cls = {"GCN": X.GCNGenerator, "GIN": X.GINGenerator, "GAT": X.GATGenerator}[gnn]
    mod = _load_params(cls(hidden, n_layers), p, "generator.").to(dev()).train(bool(training))
    visn, _, _ = O.make_inputs(seed + 1, B, 36, hidden)
    nh = {"GCN": 3, "GIN": 2, "GAT": 1}[gnn]
    masks = []
    if training:
        masks = [m for layer in O.make_keeps(seed + 3, n_layers, nh, (B, 36, hidden)) for m in layer]
    x = visn.clone().to(dev()).requires_grad_(True)
    adj = _t(gold["adj_in"]).to(dev()).requires_grad_(True)
    with inject_keep_masks(masks):
        xo, ao = mod(x, adj)
    _close(xo, gold["x_out"], name="x_out")
    _close(ao, gold["adj_out"], name="adj_out")
    assert float(torch.diagonal(ao, dim1=1, dim2=2).abs().max()) == 0.0
    ((xo * _t(gold["cx"]).to(dev())).sum() + (ao * _t(gold["ca"]).to(dev())).sum()).backward()
    _close(x.grad, gold["gx"], 2 * TOL, "gx")
    ga = adj.grad if adj.grad is not None else torch.zeros_like(adj)
    _close(ga, gold["gadj"], 2 * TOL, "gadj")
    _param_grad_check(gold, [("generator." + k, v) for k, v in mod.named_parameters()], 3 * TOL,
                      SCALAR_TOL_TC if engine == "fp32" else None)


BRANCH_CASES = [("branch_relation_gcn_h64", "relation", "GCN"), ("branch_node_gcn_h64", "node", "GCN"),
                ("branch_node_gcn_h768", "node", "GCN"), ("branch_relation_gin_h64", "relation", "GIN"),
                ("branch_relation_gcn_h768", "relation", "GCN")]


@pytest.mark.parametrize("name,which,gnn", BRANCH_CASES)
def test_ggm_branch_matches_reference_fixture(name, which, gnn, engine):
    import xggm_b200 as X
    from xggm_b200.functional import inject_keep_masks
    gold = load_golden(name)
    seed, hidden, B, n_layers, _ = [int(v) for v in gold["meta"]]
    sigma, A = float(gold["sigma"]), int(gold["num_answers"])
    p = O.make_params(seed, gnn, hidden, n_layers, 36, heads=True)
    mod = _load_params(X.XGGMHeads(hidden, gnn, n_layers), p).to(dev()).train()
    visn, xp, adj_true = O.make_inputs(seed + 1, B, 36, hidden)
    nh = {"GCN": 3, "GIN": 2}[gnn]
    masks = [m for layer in O.make_keeps(seed + 3, n_layers, nh, (B, 36, hidden)) for m in layer]
    x = xp.clone().to(dev()).requires_grad_(True)
    feat = visn.clone().to(dev()).requires_grad_(True)
    randn = _t(gold["randn"]).to(dev())
    with inject_keep_masks(masks):
        if which == "relation":
            x_gen, loss_sm, nodes, adj_g = mod.relation_step(x, feat, adj_true.to(dev()), sigma, A, 8.0, randn)
        else:
            x_gen, loss_sm, nodes, adj_g = mod.node_step(x, feat, adj_true.to(dev()), sigma, A, randn)
    _close(x_gen, gold["x_gen"], name="x_gen")
    _close(nodes, gold["nodes"], name="nodes")
    _close(adj_g, gold["adj_gen"], name="adj_gen")
    assert abs(float(loss_sm) - float(gold["loss_sm"])) < TOL * abs(float(gold["loss_sm"]))
    ((x_gen * _t(gold["c"]).to(dev())).sum() + loss_sm).backward()
    _close(x.grad, gold["gxp"], 3 * TOL, "gxp")
    gv = feat.grad if feat.grad is not None else torch.zeros_like(feat)
    _close(gv, gold["gvisn"], 3 * TOL, "gvisn")
    _param_grad_check(gold, list(mod.named_parameters()), 5 * TOL, SCALAR_TOL_TC if engine == "fp32" else None)


@pytest.mark.parametrize("gnn,B,N,H", [("GCN", 5, 36, 768), ("GCN", 2, 64, 128), ("GIN", 3, 100, 64), ("GCN", 1, 36, 768)])
def test_generator_vs_oracle_seeded(gnn, B, N, H):
    """Sizes the fixtures do not cover (other N, ragged B) against the fp64 oracle."""
    import xggm_b200 as X
    from xggm_b200.functional import inject_keep_masks
    p = O.make_params(77, gnn, H, 2, N, heads=False)
    cls = {"GCN": X.GCNGenerator, "GIN": X.GINGenerator}[gnn]
    mod = _load_params(cls(H, 2), p, "generator.").to(dev()).train()
    visn, _, adj_true = O.make_inputs(78, B, N, H)
    adj = O.strip_diag(adj_true)
    nh = {"GCN": 3, "GIN": 2}[gnn]
    keeps = O.make_keeps(79, 2, nh, (B, N, H))
    x = visn.clone().to(dev()).requires_grad_(True)
    a = adj.clone().to(dev()).requires_grad_(True)
    with inject_keep_masks([m for layer in keeps for m in layer]):
        xo, ao = mod(x, a)
    (xo.sum() + (ao * ao).sum()).backward()
    p64 = {k: v.double() for k, v in p.items()}
    x64 = visn.double().requires_grad_(True)
    a64 = adj.double().requires_grad_(True)
    xr, ar = O.GENERATORS[gnn](x64, a64, p64, 2, keeps, pre="generator.")
    (xr.sum() + (ar * ar).sum()).backward()
    _close(xo, xr, name="x")
    _close(ao, ar, name="adj")
    _close(x.grad, x64.grad, 2 * TOL, "gx")
    _close(a.grad, a64.grad, 2 * TOL, "gadj")


def test_empty_batch_and_errors():
    import xggm_b200 as X
    mod = X.GCNGenerator(64, 1).to(dev())
    x = torch.zeros(0, 36, 64, device=dev())
    adj = torch.zeros(0, 36, 36, device=dev())
    xo, ao = mod(x, adj)
    assert xo.shape == (0, 36, 64) and ao.shape == (0, 36, 36)
    with pytest.raises(RuntimeError):
        mod(torch.zeros(2, 36, 64), torch.zeros(2, 36, 36))            # CPU tensors: no fallback
    with pytest.raises(RuntimeError):
        mod(torch.zeros(2, 36, 64, device=dev()).double(), torch.zeros(2, 36, 36, device=dev()).double())
    with pytest.raises(RuntimeError):
        mod(torch.zeros(2, 36, 64, device=dev()), torch.zeros(2, 35, 35, device=dev()))
    with pytest.raises(Exception):
        X.GATGenerator(64, 2).to(dev())(torch.randn(2, 36, 64, device=dev()), torch.ones(2, 36, 36, device=dev()))


def test_full_size_properties():
    """BASELINE cfg-2 size (B=256, N=36, H=768): size-independent properties of the block."""
    import xggm_b200 as X
    torch.manual_seed(9595)
    B, N, H = 256, 36, 768
    mod = X.XGGMHeads(H, "GCN", 2).to(dev()).train()
    visn, xp, adj_true = (t.to(dev()) for t in O.make_inputs(1, B, N, H))
    x = xp.clone().requires_grad_(True)
    feat = visn.clone().requires_grad_(True)
    x_gen, loss_sm, nodes, adj_g = mod.node_step(x, feat, adj_true, 1.0, 2274)
    (x_gen.sum() + loss_sm).backward()
    assert torch.isfinite(loss_sm) and torch.isfinite(nodes).all() and torch.isfinite(adj_g).all()
    assert float(torch.diagonal(adj_g, dim1=1, dim2=2).abs().max()) == 0.0
    off = adj_g[:, ~torch.eye(N, dtype=torch.bool, device=dev())]
    assert float(off.min()) > 0.0 and float(off.max()) <= 0.7311  # sigmoid(S/colmax) <= sigmoid(1)
    # per-sample independence: graph b's output does not depend on the other graphs
    mod.eval()
    with torch.no_grad():
        full, _ = mod.generator(visn, X.glue.strip_diag(adj_true))
        part, _ = mod.generator(visn[100:103], X.glue.strip_diag(adj_true)[100:103])
    # (to the engine's rounding level ~5e-6 per product, not bitwise: the tensor-core message-passing / Gram tiles hold 3 graphs, so a
    # graph's position inside its tile -- hence the summation grouping of its dot products -- depends on b)
    assert rel_l2(part.cpu(), full[100:103].cpu()) < 2e-5 and rel_max(part.cpu(), full[100:103].cpu()) < 2e-4
    # linearity of the backward pass in the cotangent (eval mode: deterministic)
    xa = visn[:8].clone().requires_grad_(True)
    out, _ = mod.generator(xa, X.glue.strip_diag(adj_true)[:8])
    c1, c2 = torch.randn_like(out), torch.randn_like(out)
    g1, = torch.autograd.grad(out, xa, c1, retain_graph=True)
    g2, = torch.autograd.grad(out, xa, c2, retain_graph=True)
    g12, = torch.autograd.grad(out, xa, c1 + 2 * c2)
    _close(g12, (g1 + 2 * g2).cpu(), 1e-5, "linearity")
    for p_ in mod.parameters():
        if p_.grad is not None:
            assert torch.isfinite(p_.grad).all()


@pytest.mark.parametrize("B,N,gnn,precision", [(4096, 36, "GCN", "fp32"), (2048, 64, "GCN", "fp32"), (1024, 100, "GCN", "fp32"),
                                               (4096, 36, "GCN", "bf16"), (4096, 36, "GIN", "fp32"), (1024, 100, "GIN", "bf16")])
def test_max_size_properties(B, N, gnn, precision):
    """BASELINE configs[3] corner sizes (B = 4096 at obj36, 2048 x 64, 1024 x 100 nodes; H = 768: up to 147,456 node rows,
    1,152 row tiles): what is pinned to the fp64 oracle at B <= 256 must hold unchanged at the largest batch.

      * per-graph independence, forward AND backward: graphs taken from the first, a middle and the LAST tile of the big batch
        give the same outputs and input gradients as the same graphs evaluated alone (small batches are oracle-pinned),
        and graphs that receive no cotangent receive exactly zero input gradient;
      * bit-exact structure of the regenerated adjacency (zero diagonal, sigmoid range);
      * a full training step (dropout, losses, every parameter gradient) stays finite."""
    import xggm_b200 as X
    X.set_precision(precision)
    try:
        _max_size_body(X, B, N, gnn, 1.0 if precision == "fp32" else 100.0)   # bf16 engine: 2e-2 budget (a rounding flip of one
    finally:                                                                # stored bf16 value is 4e-3 of that element)
        X.set_precision("fp32")


def _max_size_body(X, B, N, gnn, slack):
    torch.manual_seed(4242)
    H = 768
    mod = X.XGGMHeads(H, gnn, 2, N).to(dev())
    g = torch.Generator().manual_seed(77)
    visn = torch.randn(B, N, H, generator=g).to(dev())
    xp = torch.randn(B, H, generator=g).to(dev())
    c = torch.rand(B, N, N, generator=g) * 1.2 - 0.2
    c = c + c.transpose(1, 2)
    adj_true = (c / c.amax(dim=(1, 2), keepdim=True)).to(dev())
    adj_in = X.glue.strip_diag(adj_true)
    mod.eval()
    picks = [0, 1, B // 2 + 1, B - 2, B - 1]          # first tile, a middle tile, the ragged last tile
    xa = visn.clone().requires_grad_(True)
    full, adj_full = mod.generator(xa, adj_in)
    cot = torch.zeros_like(full)
    cs = torch.randn(len(picks), N, H, generator=g).to(dev())
    cot[picks] = cs
    gfull, = torch.autograd.grad(full, xa, cot)
    assert float(torch.diagonal(adj_full.detach(), dim1=1, dim2=2).abs().max()) == 0.0
    off = adj_full.detach()[:, ~torch.eye(N, dtype=torch.bool, device=dev())]
    assert float(off.min()) > 0.0 and float(off.max()) <= 0.7311
    xs = visn[picks].clone().requires_grad_(True)
    part, adj_part = mod.generator(xs, adj_in[picks])
    gpart, = torch.autograd.grad(part, xs, cs)
    assert rel_l2(full[picks].detach().cpu(), part.detach().cpu()) < 2e-5 * slack and rel_max(full[picks].detach().cpu(), part.detach().cpu()) < 2e-4 * slack
    assert rel_l2(adj_full[picks].detach().cpu(), adj_part.detach().cpu()) < 2e-5 * slack
    assert rel_l2(gfull[picks].cpu(), gpart.cpu()) < 5e-5 * slack and rel_max(gfull[picks].cpu(), gpart.cpu()) < 5e-4 * slack
    rest = torch.ones(B, dtype=torch.bool)
    rest[picks] = False
    assert float(gfull[rest.to(dev())].abs().max()) == 0.0          # no cotangent -> exactly no gradient
    del full, adj_full, gfull, cot, xa
    # one full training step of the node branch at this size
    mod.train()
    x = xp.clone().requires_grad_(True)
    feat = visn.clone().requires_grad_(True)
    x_gen, loss_sm, nodes, adj_g = mod.node_step(x, feat, adj_true, 1.0, 2274)
    (x_gen.sum() + loss_sm).backward()
    assert torch.isfinite(loss_sm) and torch.isfinite(x_gen).all() and torch.isfinite(feat.grad).all() and torch.isfinite(x.grad).all()
    for p_ in mod.parameters():
        if p_.grad is not None:
            assert torch.isfinite(p_.grad).all()


@pytest.mark.parametrize("branch", ["node", "relation"])
def test_full_size_tensor_core_engine_matches_exact_engine(branch):
    """BASELINE size (B=256, N=36, H=768, train mode): the tensor-core engine -- CTA-pair kernel, grouped launches,
    K-concatenated dgrad, tensor-core message passing / Gram tiles, operand-plane hand-over; none of which the
    small fixtures reach (M <= 128 rows there) -- against the exact-fp32 engine (itself pinned to the oracle by the
    fixture tests) on the same weights, inputs, noise and Philox dropout sites: outputs, input gradients and every
    parameter gradient within the fp32 budget."""
    import xggm_b200 as X
    import xggm_b200.functional as XF
    B, N, H, A = 256, 36, 768, 2274
    torch.manual_seed(31)
    mod = X.XGGMHeads(H, "GCN", 2).to(dev()).train()
    visn, xp, adj_true = (t.to(dev()) for t in O.make_inputs(5, B, N, H))
    g = torch.Generator().manual_seed(6)
    randn = (torch.randn(B, N, H, generator=g) if branch == "node" else torch.randn(B, N, N, generator=g)).to(dev())
    cot = torch.randn(B, H, generator=g).to(dev())
    res = {}
    for engine in ("fp32", "fp32_simt"):
        X.set_precision(engine)
        try:
            mod.zero_grad(set_to_none=True)
            XF._drop.site = 500
            x = xp.clone().requires_grad_(True)
            feat = visn.clone().requires_grad_(True)
            if branch == "node":
                x_gen, loss_sm, nodes, adj_g = mod.node_step(x, feat, adj_true, 1.0, A, randn)
                ((x_gen * cot).sum() + 1.1 * loss_sm).backward()
            else:
                x_gen, loss_sm, nodes, adj_g = mod.relation_step(x, feat, adj_true, 1.0, A, kl_weight=12.0, randn=randn)
                ((x_gen * cot).sum() + 6.0 * loss_sm).backward()
            res[engine] = {"x_gen": x_gen.detach().cpu(), "loss_sm": loss_sm.detach().cpu().reshape(1),
                           "nodes": nodes.detach().cpu(), "adj": adj_g.detach().cpu(), "gx": x.grad.cpu(),
                           "gfeat": feat.grad.cpu(),
                           **{"g/" + n_: p_.grad.detach().cpu().clone() for n_, p_ in mod.named_parameters()
                              if p_.grad is not None}}
        finally:
            X.set_precision("fp32")
    assert res["fp32"].keys() == res["fp32_simt"].keys() and len(res["fp32"]) > 20
    for k, ref in res["fp32_simt"].items():
        # parameter gradients are sums over 9216 rows of products of O(1e-5)-accurate factors: 2e-4 as in the
        # fixture tests; activations and input gradients 1e-4
        _close(res["fp32"][k], ref, 2e-4 if k.startswith("g/") else TOL, f"{branch}: {k}")


def test_keep_mask_statistics_and_determinism():
    import xggm_b200.functional as XF
    torch.manual_seed(123)
    m1 = XF.keep_mask((64, 36, 768), 0.5, dev())
    m2 = XF.keep_mask((64, 36, 768), 0.5, dev())
    frac = float(m1.float().mean())
    assert abs(frac - 0.5) < 2e-3
    assert not torch.equal(m1, m2)
    m3 = XF.keep_mask((1000, 1000), 0.1, dev())
    assert abs(float(m3.float().mean()) - 0.9) < 2e-3
    assert set(m1.unique().tolist()) <= {0, 1}


def test_philox_heads_match_materialised_masks():
    """In-kernel Philox dropout (no mask tensors) == the same layer fed the masks xggm_keep_mask
    materialises for the same (seed, subsequence)."""
    import xggm_b200 as X
    import xggm_b200.functional as XF
    torch.manual_seed(1234)
    mod = X.GCNGenerator(768, 1).to(dev()).train()
    x = torch.randn(3, 36, 768, device=dev())
    adj = torch.rand(3, 36, 36, device=dev())
    XF._drop.site = 100
    xo1, ao1 = mod(x, adj)
    XF._drop.site = 100
    masks = [XF.keep_mask(x.shape, 0.5, x.device) for _ in range(3)]   # sites 100, 101, 102
    with XF.inject_keep_masks(masks):
        xo2, ao2 = mod(x, adj)
    assert torch.equal(xo1, xo2) and torch.equal(ao1, ao2)
    frac = float(torch.stack(masks).float().mean())
    assert abs(frac - 0.5) < 5e-3


def test_operand_plane_handover_is_exact_and_drops_stale_planes():
    """Planes emitted by the producer of a tensor (feature noise, previous GNN layer) and handed to its
    consumers give bit-identical results to consumers that split the tensor themselves; an in-place
    update of the tensor invalidates the hand-over."""
    import xggm_b200 as X
    import xggm_b200.functional as XF
    torch.manual_seed(5)
    mod = X.GCNGenerator(768, 2).to(dev()).eval()
    f = torch.randn(3, 768, device=dev())
    rn = torch.randn(3, 36, 768, device=dev())
    adj = torch.rand(3, 36, 36, device=dev())
    noisy, _ = X.add_feature_noise(f, sigma=1.0, randn=rn)
    assert XF._planes_of(noisy) is not None
    x1, a1 = mod(noisy, adj)
    assert XF._planes_of(x1) is not None          # emitted by the last read-out accumulation
    plain = noisy.clone()                          # a fresh tensor carries no planes
    assert XF._planes_of(plain) is None
    x2, a2 = mod(plain, adj)
    assert torch.equal(x1, x2) and torch.equal(a1, a2)
    a_direct = XF.adj_regen(x1.clone())
    assert torch.equal(a_direct, a1)
    noisy.mul_(2.0)                                # stale planes must not be used
    assert XF._planes_of(noisy) is None
    x3, _ = mod(noisy, adj)
    x4, _ = mod(plain * 2.0, adj)
    assert torch.equal(x3, x4)


def test_graphed_step_matches_eager_and_redraws_dropout():
    import xggm_b200 as X
    from xggm_b200.ddp import FlatGrads
    from xggm_b200.graphs import GraphedStep
    torch.manual_seed(7)
    B, H = 8, 768
    model = X.XGGMHeads(H, "GCN", 2).to(dev())
    grads = FlatGrads(model.parameters())
    visn, xp, adj_true = (t.to(dev()) for t in O.make_inputs(11, B, 36, H))
    randn = torch.randn(B, 36, H, device=dev())
    cot = torch.randn(B, H, device=dev())
    gx = torch.zeros(B, H, device=dev())

    def step(visn, xp, adj, randn):
        grads.zero_()
        x = xp.requires_grad_(True)
        x_gen, loss_sm, _, _ = model.node_step(x, visn, adj, 1.0, 2274, randn)
        ((x_gen * cot).sum() + 1.1 * loss_sm).backward()
        gx.copy_(x.grad)
        return x_gen.detach(), loss_sm.detach()

    model.eval()   # no dropout: graph replay must reproduce the eager step bit for bit
    xg_e, ls_e = step(visn, xp, adj_true, randn)
    flat_e, gx_e = grads.flat.clone(), gx.clone()
    g = GraphedStep(step, [visn, xp, adj_true, randn])
    xg_g, ls_g = g(visn, xp, adj_true, randn)
    torch.cuda.synchronize()
    assert torch.equal(xg_g, xg_e)
    # loss reductions and split-K weight gradients use fp32 atomics: order-dependent in the last bits only
    assert abs(float(ls_g) - float(ls_e)) <= 1e-6 * abs(float(ls_e))
    assert rel_l2(gx.cpu(), gx_e.cpu()) < 1e-6
    assert rel_l2(grads.flat.cpu(), flat_e.cpu()) < 1e-6

    model.train()  # dropout on: every replay must draw new masks, yet stay finite and reasonable
    g2 = GraphedStep(step, [visn, xp, adj_true, randn])
    a = g2.replay()[0].clone()
    b = g2.replay()[0].clone()
    torch.cuda.synchronize()
    assert torch.isfinite(a).all() and torch.isfinite(b).all()
    assert not torch.equal(a, b)


def test_fused_grad_accumulation_matches_autograd_accumulation():
    """Kernels accumulating straight into pre-existing .grad buffers == autograd's AccumulateGrad path,
    including accumulation over two backward passes."""
    import xggm_b200 as X
    import xggm_b200.functional as XF
    from xggm_b200.ddp import FlatGrads
    torch.manual_seed(3)
    B, H = 4, 768
    visn, xp, adj_true = (t.to(dev()) for t in O.make_inputs(21, B, 36, H))
    randn = torch.randn(B, 36, H, device=dev())
    out = {}
    for fused in (True, False):
        torch.manual_seed(5)
        model = X.XGGMHeads(H, "GIN", 2).to(dev()).eval()
        grads = FlatGrads(model.parameters())
        XF.FUSE_GRAD_ACCUMULATION = fused
        try:
            for _ in range(2):   # two passes: the second must ADD to the first
                x = xp.clone().requires_grad_(True)
                x_gen, loss_sm, _, _ = model.relation_step(x, visn, adj_true, 1.0, 2274, 12.0, randn[:, :, :36].contiguous())
                (x_gen.sum() + loss_sm).backward()
        finally:
            XF.FUSE_GRAD_ACCUMULATION = True
        out[fused] = grads.flat.clone()
    assert float(out[False].abs().max()) > 0
    assert rel_l2(out[True].cpu(), out[False].cpu()) < 2e-6


@pytest.mark.parametrize("gnn", ["GCN", "GIN"])
def test_bf16_engine_within_bf16_budget(gnn):
    """Single-pass bf16 tensor-core engine (BASELINE cfg 3) vs the fp64 oracle: 2e-2 on activations
    and gradients (north_star tolerance for bf16)."""
    import xggm_b200 as X
    from xggm_b200.functional import inject_keep_masks
    B, N, H = 6, 36, 768
    p = O.make_params(91, gnn, H, 2, N, heads=False)
    cls = {"GCN": X.GCNGenerator, "GIN": X.GINGenerator}[gnn]
    mod = _load_params(cls(H, 2), p, "generator.").to(dev()).train()
    visn, _, adj_true = O.make_inputs(92, B, N, H)
    adj = O.strip_diag(adj_true)
    nh = {"GCN": 3, "GIN": 2}[gnn]
    keeps = O.make_keeps(93, 2, nh, (B, N, H))
    x = visn.clone().to(dev()).requires_grad_(True)
    a = adj.clone().to(dev()).requires_grad_(True)
    X.set_precision("bf16")
    try:
        with inject_keep_masks([m for layer in keeps for m in layer]):
            xo, ao = mod(x, a)
        g = torch.Generator().manual_seed(94)
        cx, ca = torch.randn(B, N, H, generator=g), torch.randn(B, N, N, generator=g)
        ((xo * cx.to(dev())).sum() + (ao * ca.to(dev())).sum()).backward()
    finally:
        X.set_precision("fp32")
    p64 = {k: v.double().requires_grad_(True) for k, v in p.items()}
    x64, a64 = visn.double().requires_grad_(True), adj.double().requires_grad_(True)
    fn = O.gcn_generator if gnn == "GCN" else O.gin_generator
    xr, ar = fn(x64, a64, p64, 2, keeps, pre="generator.")
    ((xr * cx.double()).sum() + (ar * ca.double()).sum()).backward()
    BF16_TOL = 2e-2
    assert rel_l2(xo.detach().cpu(), xr.detach()) < BF16_TOL
    assert rel_l2(ao.detach().cpu(), ar.detach()) < BF16_TOL
    assert float(torch.diagonal(ao, dim1=1, dim2=2).abs().max()) == 0.0   # masks stay bit-exact
    assert rel_l2(x.grad.cpu(), x64.grad) < 2 * BF16_TOL
    assert rel_l2(a.grad.cpu(), a64.grad) < 2 * BF16_TOL


def test_api_surface_compositions_discriminator_v2_and_mix_generator():
    """ggm.py:85-97 and :272-323 (never built by the trainers): thin compositions over the library kernels,
    checked against the same composition written with stock torch ops in fp64."""
    import torch.nn.functional as F
    import xggm_b200 as X
    torch.manual_seed(17)
    H, B = 128, 5
    d2 = X.DiscriminatorV2(36 * H).to(dev())
    g = torch.randn(B, 36, H, device=dev())
    got = d2(g)
    m = d2.model.double()
    ref = m(g.double().view(B, -1))
    d2.float()
    _close(got, ref.detach().cpu(), name="DiscriminatorV2")

    mix = X.MixGenerator(H, 2).to(dev()).eval()
    x = torch.randn(B, H, device=dev())
    adj = torch.rand(B, 36, 36, device=dev())
    obj = torch.rand(B, 36, H, device=dev())
    eps = torch.randn(B, H, device=dev())
    nodes, loss = mix(x, adj, obj, eps)
    # stock-torch fp64 restatement of ggm.py:296-323 + gin.py:21-34,68-87 with the same weights
    p = {k: v.detach().double().cpu() for k, v in mix.state_dict().items()}
    xd, ad, od, ed = x.double().cpu(), adj.double().cpu(), obj.double().cpu(), eps.double().cpu()
    mu = F.linear(xd, p["fc1.weight"], p["fc1.bias"]); lv = F.linear(xd, p["fc2.weight"], p["fc2.bias"])
    z = mu + lv.mul(0.5).exp() * ed
    h = F.linear(z, p["decoder.0.weight"], p["decoder.0.bias"])
    h = torch.relu(F.layer_norm(h, (6 * H,), p["decoder.1.weight"], p["decoder.1.bias"]))
    nf = F.linear(h, p["decoder.3.weight"], p["decoder.3.bias"]).view(-1, 36, H)
    ref_loss = F.binary_cross_entropy_with_logits(nf, od) * 768 - 0.5 * torch.sum(1 + lv - mu.pow(2) - lv.exp())
    for l in range(2):
        nf = O.gin(nf, ad, p, f"gnn_layers.{l}.", 1, None)
    _close(nodes, nf, 2 * TOL, "MixGenerator nodes")
    assert abs(float(loss) - float(ref_loss)) < 1e-4 * abs(float(ref_loss))


@pytest.mark.parametrize("B,N,H", [(7, 36, 768), (5, 64, 128), (3, 100, 64), (4, 36, 776)])
def test_tensor_core_graph_ops_match_simt_and_stay_in_bounds(B, N, H):
    """Tensor-core message passing (block-diagonal tiles) and Gram regeneration vs the SIMT kernels on
    ragged graph counts (B not a multiple of the graphs-per-tile) with guard bands around the outputs."""
    import xggm_b200.functional as XF
    from xggm_b200._lib import call, ptr
    d = dev()
    g = torch.Generator().manual_seed(B * 1000 + N)
    x = torch.randn(B, N, H, generator=g).to(d)
    adj = torch.rand(B, N, N, generator=g).to(d)
    GUARD, SENT = 2048, 777.0
    n = B * N * H
    res = {}
    for tag, wk in (("simt", None), ("tc", XF._adj_work(B, N, H, d))):
        buf = torch.full((n + 2 * GUARD,), SENT, device=d)
        out = buf[GUARD:GUARD + n]
        call("xggm_adj_apply_fwd", ptr(adj), ptr(x), ptr(out), B, N, H, 0.5, None, 1.0, ptr(wk))
        gx = torch.ones(n, device=d)
        call("xggm_adj_apply_bwd", ptr(adj), ptr(x), ptr(x), ptr(gx), None, B, N, H, 1.0, None, 0.0, 1, ptr(wk))
        torch.cuda.synchronize()
        assert bool((buf[:GUARD] == SENT).all()) and bool((buf[GUARD + n:] == SENT).all()), tag
        res[tag] = (out.clone(), gx.clone())
    ref_f = x.double() + 0.5 * torch.bmm(adj.double(), x.double())
    ref_b = 1.0 + torch.bmm(adj.double().transpose(1, 2), x.double())
    for tag in ("simt", "tc"):
        assert rel_l2(res[tag][0].cpu(), ref_f.reshape(-1).cpu()) < 3e-5, tag
        assert rel_l2(res[tag][1].cpu(), ref_b.reshape(-1).cpu()) < 3e-5, tag
    a1, am1 = XF.adj_regen(x, True, return_argmax=True)
    s = torch.bmm(x.double(), x.double().transpose(1, 2))
    ref = torch.sigmoid(s / s.max(dim=1)[0].unsqueeze(-1))
    ref = ref - torch.diag_embed(torch.diagonal(ref, dim1=1, dim2=2))
    assert rel_l2(a1.cpu(), ref.cpu()) < 1e-4
    assert float(torch.diagonal(a1, dim1=1, dim2=2).abs().max()) == 0.0


@pytest.mark.parametrize("name", ["visual_feat_train", "visual_feat_eval"])
def test_visual_feat_encoder_matches_reference_fixture(name, engine):
    """xggm_b200.VisualFeatEncoder vs the fixture written by lxrt.modeling.VisualFeatEncoder
    (src/lxrt/modeling.py:530-556); state_dict keys are the reference's."""
    import xggm_b200 as X
    from xggm_b200.functional import inject_keep_masks
    gold = load_golden(name)
    seed, hidden, B, training = [int(v) for v in gold["meta"]]
    mod = X.VisualFeatEncoder(hidden_size=hidden, hidden_dropout_prob=float(gold["drop_p"]))
    mod.load_state_dict(O.make_visual_params(seed, hidden), strict=True)
    mod = mod.to(dev()).train(bool(training))
    feats, boxes = O.make_visual_inputs(seed + 1, B)
    f = feats.clone().to(dev()).requires_grad_(True)
    b = boxes.clone().to(dev()).requires_grad_(True)
    masks = [_t(gold["keep"])] if training else []
    with inject_keep_masks(masks):
        out = mod((f, b))
    _close(out, gold["out"], name="out")
    if training:   # dropped positions are exactly zero
        assert float(out[_t(gold["keep"]).to(dev()) == 0].abs().max()) == 0.0
    (out * _t(gold["c"]).to(dev())).sum().backward()
    nrm = float(gold["gfeats_norm"][0])
    assert abs(float(f.grad.double().norm()) - nrm) < 3 * TOL * nrm
    _close(f.grad.reshape(-1)[:64], gold["gfeats_head"], 20 * TOL, "gfeats")
    _close(b.grad, gold["gboxes"], 3 * TOL, "gboxes")
    named = dict(mod.named_parameters())
    for k in ("box_fc.weight", "box_fc.bias", "visn_layer_norm.weight", "box_layer_norm.bias"):
        _close(named[k].grad, gold["g/" + k], 3 * TOL, k)
    _param_grad_check(gold, list(mod.named_parameters()), 3 * TOL)


# --------------------------------------------------------------------------- training-step tail
def test_bertadam_clip_and_bce_match_reference_fixture():
    """xggm_grad_sumsq + xggm_bertadam_step (clip applied on the fly) and xggm_bce_logits_* against the
    reference BertAdam / clip_grad_norm_ / BCEWithLogitsLoss fixture (tests/golden/optimizer.npz)."""
    import xggm_b200 as X
    g = load_golden("optimizer")
    steps = int(g["seed"][1])
    n = sum(1 for k in g if k.startswith("p0/"))
    params = [torch.nn.Parameter(_t(g[f"p0/{i}"]).to(dev())) for i in range(n)]
    opt = X.BertAdam(params, lr=float(g["lr"][0]), warmup=float(g["warmup"][0]), t_total=int(g["t_total"][0]),
                     weight_decay=float(g["weight_decay"][0]))
    grp = opt.groups[0]
    for s in range(steps):
        for i, p in enumerate(params):
            p.grad.copy_(_t(g[f"g{s}/{i}"]))
        clip = X.clip_grad_norm_(opt, float(g["max_norm"][0]))
        assert abs(clip.total_norm() - float(g[f"norm{s}"][0])) <= 1e-5 * float(g[f"norm{s}"][0])
        opt.step(clip)
        for i, p in enumerate(params):
            _close(p, g[f"p{s + 1}/{i}"], 1e-6, f"p step {s}")
            o = (p.data_ptr() - grp.flat_p.data_ptr()) // 4
            _close(grp.m[o:o + p.numel()].view_as(p), g[f"m{s + 1}/{i}"], 1e-6, f"m step {s}")
            _close(grp.v[o:o + p.numel()].view_as(p), g[f"v{s + 1}/{i}"], 2e-6, f"v step {s}")
    x = _t(g["bce/logit"]).to(dev()).requires_grad_(True)
    loss = X.bce_with_logits(x, _t(g["bce/target"]).to(dev()), scale=x.shape[1])
    (1.0 * loss).backward()
    assert abs(float(loss) - float(g["bce/loss"][0])) <= 1e-5 * abs(float(g["bce/loss"][0]))
    _close(x.grad, g["bce/glogit"], 1e-5, "glogit")


def test_bertadam_full_size_properties():
    """At the block's real size (8.2 M parameters): zero gradients and zero weight decay leave the parameters
    bit-identical; the update is elementwise (a permuted problem gives the permuted result); padding between
    tensors stays zero."""
    import xggm_b200 as X
    torch.manual_seed(3)
    model = X.XGGMHeads(768, "GCN", 2, 36).to(dev())
    opt = X.BertAdam(model.parameters(), lr=1e-3, weight_decay=0.0)
    grp = opt.groups[0]
    before = grp.flat_p.clone()
    opt.step(X.clip_grad_norm_(opt, 5.0))
    assert torch.equal(grp.flat_p, before) and float(grp.m.abs().max()) == 0.0
    grp.grads.flat.normal_()
    mask = torch.zeros_like(grp.flat_p, dtype=torch.bool)
    for p in grp.params:
        o = (p.data_ptr() - grp.flat_p.data_ptr()) // 4
        mask[o:o + p.numel()] = True
    grp.grads.flat.mul_(mask)
    clip = X.clip_grad_norm_(opt, 5.0)
    ref_norm = float(grp.grads.flat.double().norm())
    assert abs(clip.total_norm() - ref_norm) <= 1e-5 * ref_norm
    opt.step(clip)
    if bool((~mask).any()):
        assert float(grp.flat_p[~mask].abs().max()) == 0.0
    coef = min(1.0, 5.0 / (ref_norm + 1e-6))
    gq = grp.grads.flat.double() * coef
    want = before.double() - 1e-3 * ((0.1 * gq) / ((0.001 * gq * gq).sqrt() + 1e-6))
    _close(grp.flat_p, want.float().cpu(), 1e-6, "one step from zero moments")


def test_training_iteration_node_branch_matches_oracle_pipeline():
    """Three complete GGM training iterations of the node branch the way the trainer runs them
    (src/vqa/vqacpv2.py:226-254): node_step -> logit_fc -> BCEWithLogits * A + 1.1 * loss_sm -> backward ->
    clip_grad_norm_(5.) -> BertAdam.step, through the library (flat buffers, fused kernels) and through the fp64
    oracle pipeline on the same weights, inputs, noise and dropout masks.  Checks the losses, the clip norm and
    the parameter UPDATES of both iterations."""
    import xggm_b200 as X
    from xggm_b200.functional import inject_keep_masks
    B, N, H, L, A, sigma, lr = 3, 36, 128, 2, 50, 1.0, 1e-3
    p = O.make_params(11, "GCN", H, L, N, heads=True)
    g = torch.Generator().manual_seed(12)
    head = {"logit_fc.0.weight": torch.randn(2 * H, H, generator=g) * 0.02, "logit_fc.0.bias": torch.zeros(2 * H),
            "logit_fc.2.weight": torch.ones(2 * H), "logit_fc.2.bias": torch.zeros(2 * H),
            "logit_fc.3.weight": torch.randn(A, 2 * H, generator=g) * 0.02, "logit_fc.3.bias": torch.zeros(A)}
    mod = X.XGGMHeads(H, "GCN", L, N).to(dev()).train()
    mod.load_state_dict(p, strict=True)
    ans = X.AnswerHead(H, A).to(dev()).train()
    ans.load_state_dict(head, strict=True)
    # encoder_adj takes no part in the node branch: the trainer's optimiser skips it (its .grad stays None there);
    # here it IS registered with the optimiser, which must leave it untouched (active ranges of the bucket)
    params = [q for n_, q in mod.named_parameters() if not n_.startswith("encoder_adj")] + list(ans.parameters())
    names = [n_ for n_, _ in mod.named_parameters() if not n_.startswith("encoder_adj")] + [n_ for n_, _ in ans.named_parameters()]
    idle = [q for n_, q in mod.named_parameters() if n_.startswith("encoder_adj")]
    idle_before = [q.detach().clone() for q in idle]
    opt = X.BertAdam(params + idle, lr=lr, warmup=0.1, t_total=20)
    ref = {k: v.double().clone() for k, v in {**p, **head}.items()}
    mom = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in ref.items()}
    for it in range(3):   # (the warm-up schedule makes the first step a zero-length one: lr(0) = 0)
        visn, xp, adj_true = O.make_inputs(100 + it, B, N, H)
        keeps = O.make_keeps(200 + it, L, 3, (B, N, H))
        randn = torch.randn(B, N, H, generator=g)
        target = (torch.rand(B, A, generator=g) < 0.05).float()
        # ---- library
        opt.zero_grad()
        x = xp.clone().to(dev()).requires_grad_(True)
        feat = visn.clone().to(dev()).requires_grad_(True)
        with inject_keep_masks([m for layer in keeps for m in layer]):
            x_gen, loss_sm, _, _ = mod.node_step(x, feat, adj_true.to(dev()), sigma, A, randn.to(dev()))
        loss = X.bce_with_logits(ans(x_gen), target.to(dev()), scale=A) + 1.1 * loss_sm
        loss.backward()
        clip = X.clip_grad_norm_(opt, 5.0)
        before = {n_: q.detach().clone() for n_, q in zip(names, params)}
        opt.step(clip)
        # ---- oracle (fp64)
        rp = {k: v.clone().requires_grad_(True) for k, v in ref.items()}
        xg_r, ls_r, _, _ = O.node_branch(xp.double(), visn.double(), adj_true.double(), rp, sigma, randn.double(), keeps, A)
        hdn = torch.nn.functional.layer_norm(
            O.gelu_erf(xg_r @ rp["logit_fc.0.weight"].T + rp["logit_fc.0.bias"]), (2 * H,), rp["logit_fc.2.weight"],
            rp["logit_fc.2.bias"], 1e-12)
        logit_r = hdn @ rp["logit_fc.3.weight"].T + rp["logit_fc.3.bias"]
        loss_r = O.bce_with_logits(logit_r, target.double(), scale=A) + 1.1 * ls_r
        loss_r.backward()
        live = [k for k in names if rp[k].grad is not None]
        assert set(live) == set(names)
        norm_r, coef = O.clip_coef([rp[k].grad for k in live], 5.0)
        lr_s = O.scheduled_lr(lr, it, 20, 0.1)
        assert abs(float(loss) - float(loss_r)) <= 1e-4 * abs(float(loss_r))
        assert abs(clip.total_norm() - norm_r) <= 2e-4 * norm_r
        for k, q in zip(names, params):
            new_p, m1, v1 = O.bertadam_step(ref[k], rp[k].grad * coef, mom[k][0], mom[k][1], lr_s)
            upd_ref = (new_p - ref[k])
            upd = (q.detach().cpu().double() - before[k].cpu().double())
            # the Adam direction m / (sqrt(v) + e) is sign-like where |g| is tiny: compare updates against the
            # step size lr * 3.2 (|update| <= lr * (0.1 / sqrt(0.001)) in the first steps), not element-relative
            scale_ = max(lr_s * 3.2, 1e-12)
            assert float((upd - upd_ref).abs().max()) <= 0.02 * scale_ + 1e-9, (it, k)
            assert float((upd - upd_ref).norm()) <= 2e-3 * float(upd_ref.norm()) + 1e-12, (it, k)
            ref[k], mom[k] = new_p.detach(), (m1.detach(), v1.detach())
    for q, b in zip(idle, idle_before):
        assert torch.equal(q.detach(), b)


def test_bertadam_schedule_on_device_survives_graph_replay_and_skips_unused_parameters():
    """(1) the schedule (warmup_linear, warmup 0.1, t_total) is evaluated from a DEVICE step counter the
    kernel advances, so a captured optimiser step follows it across replays instead of freezing the capture-time
    lr; (2) parameters that received no gradient since zero_grad() are skipped as the reference does
    (src/lxrt/optimization.py:139-141): no weight decay, no moment decay; (3) state_dict round trip."""
    import xggm_b200 as X
    import xggm_b200.functional as XF
    torch.manual_seed(5)
    lin_a = torch.nn.Linear(24, 16).to(dev())
    lin_b = torch.nn.Linear(16, 8).to(dev())        # takes no part in the step
    params = list(lin_a.parameters()) + list(lin_b.parameters())
    opt = X.BertAdam(params, lr=1e-2, warmup=0.1, t_total=20, weight_decay=0.01)
    grp = opt.groups[0]
    x = torch.randn(32, 24, device=dev())
    p0 = [q.detach().clone().cpu().double() for q in params]

    def step():
        opt.zero_grad()
        XF.linear(x, lin_a.weight, lin_a.bias).square().sum().backward()
        opt.step(X.clip_grad_norm_(opt, 5.0))

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()                                        # eager step 0 (lr(0) = 0 under warm-up)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step()                                        # captured (a capture does not execute)
    for _ in range(5):
        graph.replay()                                # steps 1..5
    torch.cuda.synchronize()
    assert grp.step == 6
    # oracle: the same six steps with the host-side schedule
    ref = [t.clone() for t in p0[:2]]
    mom = [(torch.zeros_like(t), torch.zeros_like(t)) for t in ref]
    xd = x.cpu().double()
    for it in range(6):
        rp = [t.clone().requires_grad_(True) for t in ref]
        (xd @ rp[0].T + rp[1]).square().sum().backward()
        _, coef = O.clip_coef([t.grad for t in rp], 5.0)
        lr_s = O.scheduled_lr(1e-2, it, 20, 0.1)
        for i in range(2):
            new_p, m1, v1 = O.bertadam_step(ref[i], rp[i].grad * coef, mom[i][0], mom[i][1], lr_s)
            ref[i], mom[i] = new_p.detach(), (m1.detach(), v1.detach())
    for i in range(2):
        moved = float((ref[i] - p0[i]).abs().max())
        assert moved > 1e-3                            # a frozen capture-time lr (lr(1) = 5e-3 * ...) would not match
        assert float((params[i].detach().cpu().double() - ref[i]).abs().max()) < 2e-3 * moved
    for i in (2, 3):                                   # unused parameters: bit-identical, moments still zero
        assert torch.equal(params[i].detach().cpu().double(), p0[i])
        o = grp.grads.offsets[i]
        assert float(grp.m[o:o + params[i].numel()].abs().max()) == 0.0
    sd = opt.state_dict()
    assert sd["state"][0]["step"] == 6 and float(sd["state"][0]["next_v"].abs().max()) > 0


def test_prepared_weight_planes_are_exact_and_follow_the_weights():
    """Prepared weight planes (xggm_weight_planes_build; functional.weight_planes): a layer called with cached
    planes gives bit-identical outputs and gradients to one that splits its weights per call; the optimiser step
    rebuilds them; a torch-visible in-place change of a weight invalidates its record."""
    import xggm_b200 as X
    import xggm_b200.functional as XF
    B, N, H = 8, 36, 256
    torch.manual_seed(21)
    mod = X.XGGMHeads(H, "GCN", 2).to(dev()).train()
    visn, xp, adj_true = (t.to(dev()) for t in O.make_inputs(22, B, N, H))
    randn = torch.randn(B, N, H, device=dev())

    def run():
        XF._drop.site = 900
        for p_ in mod.parameters():
            p_.grad = None
        x = xp.clone().requires_grad_(True)
        feat = visn.clone().requires_grad_(True)
        x_gen, loss_sm, nodes, adj_g = mod.node_step(x, feat, adj_true, 1.0, 100, randn)
        (x_gen.sum() + loss_sm).backward()
        return [x_gen.detach().clone(), nodes.detach().clone(), x.grad.clone(), feat.grad.clone()] + \
               [p_.grad.clone() for p_ in mod.parameters() if p_.grad is not None]

    def run_rel():      # the relation branch exercises the ragged 630-wide encoder_adj weight (pitched W^T planes)
        XF._drop.site = 950
        for p_ in mod.parameters():
            p_.grad = None
        x = xp.clone().requires_grad_(True)
        feat = visn.clone().requires_grad_(True)
        x_gen, loss_sm, nodes, adj_g = mod.relation_step(x, feat, adj_true, 1.0, 100, kl_weight=8.0, randn=randn_rel)
        (x_gen.sum() + loss_sm).backward()
        return [x_gen.detach().clone(), adj_g.detach().clone(), x.grad.clone(), feat.grad.clone()] + \
               [p_.grad.clone() for p_ in mod.parameters() if p_.grad is not None]

    randn_rel = torch.randn(B, N, N, device=dev())
    plain = run()
    plain_rel = run_rel()
    mats = [p_ for p_ in mod.parameters() if p_.dim() == 2]
    XF.cache_weight_planes(mats)
    l0 = X._lib.kernel_launches()
    cached_first = run()
    l1 = X._lib.kernel_launches()
    cached_again = run()
    l2 = X._lib.kernel_launches()
    assert all(getattr(m_, "_xggm_wp", None) is not None for m_ in mats if m_ is not mod.encoder_adj[0].weight)
    assert getattr(mod.encoder_adj[0].weight, "_xggm_wp", None) is None      # (not used by the node branch)
    assert l2 - l1 < l1 - l0                      # the second call found every record valid: no build launches
    def same(u, v):
        # outputs and input gradients are deterministic: bit-identical.  Parameter gradients are reduced with
        # atomics (column sums, split-K weight gradients): equal to rounding
        for i, (a, b) in enumerate(zip(u, v)):
            if i < 4:
                assert torch.equal(a, b), i
            else:
                assert rel_l2(a.cpu(), b.cpu()) < 2e-6, i

    same(plain, cached_first)
    same(plain, cached_again)
    same(plain_rel, run_rel())
    assert getattr(mod.encoder_adj[0].weight, "_xggm_wp", None) is not None
    # a visible in-place update invalidates the record; the next call rebuilds and matches the uncached path
    w = mod.generator.gnn_layers[0].linear_prediction[0][0].weight
    with torch.no_grad():
        w.mul_(1.5)
    assert XF._wp_valid(w) is None
    after = run()
    XF.cache_weight_planes(mats, enable=False)
    after_plain = run()
    same(after, after_plain)
    assert not torch.equal(after[0], plain[0])
    # BertAdam opts its parameters in and refreshes the planes after its (version-invisible) update kernel
    opt = X.BertAdam(mod.parameters(), lr=1e-2)
    run_a = run()                                  # builds planes of the current weights
    opt.step()                                     # weights move; planes are rebuilt by the step
    moved = run()
    XF.cache_weight_planes(mats, enable=False)
    moved_plain = run()
    assert not torch.equal(moved[0], run_a[0])
    same(moved, moved_plain)


# --------------------------------------------------------------------------- BASELINE size against the fp64 oracle
def _oracle_branch_fp64(branch, p, visn, xp, adj_true, randn, keeps, cot, A, w_sm, kl_weight):
    """The fp64 CPU oracle of one GGM branch with every gradient (seconds at B=256 on the box's cores)."""
    p64 = {k: v.double().requires_grad_(True) for k, v in p.items()}
    x64, f64 = xp.double().requires_grad_(True), visn.double().requires_grad_(True)
    if branch == "node":
        xg, ls, nodes, adj_g = O.node_branch(x64, f64, adj_true.double(), p64, 1.0, randn.double(), keeps, A)
    else:
        xg, ls, nodes, adj_g = O.relation_branch(x64, f64, adj_true.double(), p64, 1.0, randn.double(), keeps, A,
                                                 kl_weight=kl_weight)
    ((xg * cot.double()).sum() + w_sm * ls).backward()
    ref = {"x_gen": xg.detach(), "loss_sm": ls.detach().reshape(1), "nodes": nodes.detach(), "adj": adj_g.detach(),
           "gx": x64.grad, "gfeat": f64.grad if f64.grad is not None else torch.zeros_like(f64)}
    for k, v in p64.items():
        if v.grad is not None:
            ref["g/" + k] = v.grad
    return ref


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("branch", ["node", "relation"])
def test_full_size_branch_matches_fp64_oracle(branch, precision):
    """BASELINE size (B=256, N=36, H=768, train mode, M = 9216 rows): the tensor-core engines -- CTA-pair kernel,
    grouped launches, K-concatenated dgrad, tensor-core message passing / Gram tiles, split-K weight gradients,
    operand-plane hand-over, none of which the reference fixtures (M <= 108 rows) reach -- against the fp64 CPU
    ORACLE (itself pinned to the reference by tests/test_oracle_golden.py) on the same weights, inputs, noise and
    injected dropout masks: outputs, input gradients and every parameter gradient.
    fp32 engine: 1e-4 activations / input gradients, 2e-4 parameter gradients (sums over 9216 rows);
    bf16 engine: 2e-2 (north_star), parameter gradients included."""
    import xggm_b200 as X
    from xggm_b200.functional import inject_keep_masks
    B, N, H, A = 256, 36, 768, 2274
    p = O.make_params(41, "GCN", H, 2, N, heads=True)
    visn, xp, adj_true = O.make_inputs(42, B, N, H)
    keeps = O.make_keeps(43, 2, 3, (B, N, H))
    g = torch.Generator().manual_seed(44)
    randn = torch.randn(B, N, H, generator=g) if branch == "node" else torch.randn(B, N, N, generator=g)
    cot = torch.randn(B, H, generator=g)
    w_sm, klw = (1.1, None) if branch == "node" else (6.0, 12.0)
    mod = _load_params(X.XGGMHeads(H, "GCN", 2), p).to(dev()).train()
    x = xp.clone().to(dev()).requires_grad_(True)
    feat = visn.clone().to(dev()).requires_grad_(True)
    X.set_precision(precision)
    try:
        with inject_keep_masks([m for layer in keeps for m in layer]):
            if branch == "node":
                x_gen, loss_sm, nodes, adj_g = mod.node_step(x, feat, adj_true.to(dev()), 1.0, A, randn.to(dev()))
            else:
                x_gen, loss_sm, nodes, adj_g = mod.relation_step(x, feat, adj_true.to(dev()), 1.0, A, kl_weight=klw,
                                                                 randn=randn.to(dev()))
        ((x_gen * cot.to(dev())).sum() + w_sm * loss_sm).backward()
        torch.cuda.synchronize()
    finally:
        X.set_precision("fp32")
    got = {"x_gen": x_gen, "loss_sm": loss_sm.reshape(1), "nodes": nodes, "adj": adj_g, "gx": x.grad,
           "gfeat": feat.grad if feat.grad is not None else torch.zeros_like(feat)}
    for k, v in mod.named_parameters():
        if v.grad is not None:
            got["g/" + k] = v.grad
    ref = _oracle_branch_fp64(branch, p, visn, xp, adj_true, randn, keeps, cot, A, w_sm, klw)
    assert got.keys() == ref.keys() and len(got) > 20
    assert float(torch.diagonal(adj_g, dim1=1, dim2=2).abs().max()) == 0.0      # bit-exact mask in both engines
    worst = {}
    for k, r in ref.items():
        a = got[k].detach().cpu().double()
        worst[k] = (rel_l2(a, r), rel_max(a, r))
    if precision == "fp32":
        for k, (e2, em) in worst.items():
            tol = 2e-4 if k.startswith("g/") else TOL
            assert e2 < tol and em < 5 * tol, f"{branch}/{precision} {k}: rel_l2={e2:.3e} rel_max={em:.3e}"
    else:
        BF16_TOL = 2e-2   # relative L2, activations AND gradients
        for k, (e2, em) in worst.items():
            assert e2 < BF16_TOL, f"{branch}/{precision} {k}: rel_l2={e2:.3e}"


# --------------------------------------------------------------------------- SURVEY 8 a-18 against the reference fixture
API_CASES = ["edge_generator", "node_generator", "gin_plain_encoder", "gcn_plain_encoder", "discriminator",
             "discriminator_v2", "gcn_conv_dropout", "gat_mean"]


@pytest.mark.parametrize("tag", API_CASES)
def test_api_surface_matches_reference_fixture(tag, engine):
    """ggm.py:15-159 (EdgeGenerator, NodeGenerator, the plain encoders, both discriminators), GCNConv(dropout>0)
    (gcn.py:28) and GAT(merge != 'cat') (gat.py:76-77): the library modules load the REFERENCE module's own
    state_dict and are compared with the outputs / gradients the reference class produced (api_surface.npz)."""
    import xggm_b200 as X
    from xggm_b200.functional import inject_keep_masks
    gold = load_golden("api_surface")
    seed, hidden, B, n_layers = [int(v) for v in gold["meta"]]
    build = {"edge_generator": lambda: X.EdgeGenerator(hidden, n_layers),
             "node_generator": lambda: X.NodeGenerator(hidden, n_layers),
             "gin_plain_encoder": lambda: X.GinPlainEncoder(hidden, n_layers),
             "gcn_plain_encoder": lambda: X.GCNPlainEncoder(hidden, n_layers),
             "discriminator": lambda: X.Discriminator(2 * hidden),
             "discriminator_v2": lambda: X.DiscriminatorV2(2 * hidden),
             "gcn_conv_dropout": lambda: X.GCNConv(hidden, dropout=0.25),
             "gat_mean": lambda: X.GAT(hidden, hidden, n_head=2, merge="mean")}[tag]
    mod = build()
    sd = {k[len(tag) + 4:]: _t(v) for k, v in gold.items() if k.startswith(tag + "/sd/")}
    mod.load_state_dict(sd, strict=True)      # the reference module's own keys and shapes
    mod = mod.to(dev()).train()
    masks = []
    while f"{tag}/keep{len(masks)}" in gold:
        masks.append(_t(gold[f"{tag}/keep{len(masks)}"]))
    x = _t(gold["x_in"]).to(dev()).requires_grad_(True)
    adj = _t(gold["adj_in"]).to(dev()).requires_grad_(True)
    with inject_keep_masks(masks):
        if tag.startswith("discriminator"):
            y = mod(x[:, :2])
        elif tag == "gat_mean":
            y = mod(x, adj).reshape(1)
        else:
            y = mod(x, adj)
    _close(y, gold[tag + "/y"], name=tag + " y")
    (y * _t(gold[tag + "/c"]).to(dev())).sum().backward()
    _close(x.grad, gold[tag + "/gx"], 3 * TOL, tag + " gx")
    ga = adj.grad if adj.grad is not None else torch.zeros_like(adj)
    _close(ga, gold[tag + "/gadj"], 3 * TOL, tag + " gadj")
    # GIN's scalar eps gradient <gpre, adj @ h> is a cancellation-heavy sum (see SCALAR_TOL_TC above): one of the
    # fixture's (node_generator layer 0) is -0.18 where its siblings are ~10 and torch's own fp32 value is already
    # 1.5e-4 from fp64.  Under the tensor-core engine scalars are therefore judged against the scale of the
    # module's scalar gradients (the size of the terms that cancel), not against their own magnitude.
    scalars = [abs(float(gold[f"{tag}/g/{k}"].reshape(-1)[0])) for k, v in mod.named_parameters() if v.numel() == 1]
    for k, v in mod.named_parameters():
        gp = v.grad if v.grad is not None else torch.zeros_like(v)
        ref = gold[f"{tag}/g/{k}"]
        if v.numel() == 1 and engine == "fp32":
            err = abs(float(gp.reshape(-1)[0]) - float(ref.reshape(-1)[0]))
            assert err <= SCALAR_TOL_TC * max(scalars), f"{tag} g/{k}: abs err {err:.3e} vs scale {max(scalars):.3e}"
        else:
            _close(gp, ref, 5 * TOL, f"{tag} g/{k}")


def test_fused_message_passing_layernorm_cluster_kernel_matches_the_two_kernel_path():
    """adj_ln_tc (XGGM_ADJ_LN_TC=1: message passing + LayerNorm in one 4-CTA-cluster kernel, u kept in TMEM, row
    statistics exchanged through distributed shared memory) against the default adj_apply_tc + layernorm pair, in a
    child process (the switch is read once per process): outputs and gradients of a GCN generator at B=40 (a ragged
    last row tile) agree to the engines' rounding."""
    import os
    import subprocess
    import sys
    code = r"""
import sys, torch
sys.path.insert(0, %r)
import xggm_b200 as X
from oracle import xggm_oracle as O
torch.manual_seed(3)
dev = torch.device('cuda:0')
mod = X.GCNGenerator(768, 2).to(dev).eval()
visn, _, adj_true = O.make_inputs(5, 40, 36, 768)
x = visn.to(dev).requires_grad_(True)
a = O.strip_diag(adj_true).to(dev).requires_grad_(True)
xo, ao = mod(x, a)
g = torch.Generator().manual_seed(1)
cx, ca = torch.randn(40, 36, 768, generator=g).to(dev), torch.randn(40, 36, 36, generator=g).to(dev)
((xo * cx).sum() + (ao * ca).sum()).backward()
torch.save({'xo': xo.detach().cpu(), 'ao': ao.detach().cpu(), 'gx': x.grad.cpu(), 'ga': a.grad.cpu(),
            'gw': mod.gnn_layers[0].gnn_layers[0].layer_norm.weight.grad.cpu()}, sys.argv[1])
""" % ROOT_DIR
    import tempfile
    outs = {}
    with tempfile.TemporaryDirectory() as d:
        for flag in ("0", "1"):
            path = os.path.join(d, f"o{flag}.pt")
            env = dict(os.environ, XGGM_ADJ_LN_TC=flag)
            subprocess.run([sys.executable, "-c", code, path], check=True, env=env, timeout=300)
            outs[flag] = torch.load(path)
    for k in outs["0"]:
        _close(outs["1"][k], outs["0"][k], 2e-5 if k in ("xo", "ao") else 1e-4, k)


def test_in_kernel_gaussian_feature_noise():
    """add_feature_noise without a randn tensor (xggm_feat_noise_philox: Philox4x32-10 + Box-Muller inside the kernel):
    standard-normal statistics, target = -noise / sigma^2 exactly, determinism under torch.manual_seed, fresh draws per
    call, the broadcast ([B,H] -> [B,N,H]) form, and operand planes equal to a split of the noisy tensor."""
    import xggm_b200 as X
    import xggm_b200.functional as XF
    B, N, H, sigma = 64, 36, 768, 0.7
    f = torch.randn(B, H, device=dev())
    torch.manual_seed(11)
    XF._drop.site = 0
    noisy, target = X.add_feature_noise(f, sigma=sigma, n_nodes=N)
    n = (noisy - f[:, None, :]).double() / sigma
    assert abs(float(n.mean())) < 3e-3 and abs(float(n.std()) - 1.0) < 3e-3
    assert abs(float((n ** 3).mean())) < 2e-2 and abs(float((n ** 4).mean()) - 3.0) < 5e-2       # skewness, kurtosis
    flat = n.reshape(-1)
    assert abs(float((flat[:-1] * flat[1:]).mean())) < 3e-3                                        # neighbours uncorrelated
    assert float((flat.abs() > 4.0).double().mean()) < 3e-4 and float(flat.abs().max()) < 7.0      # tails
    # target = -noise / sigma^2 with the noise that was actually added (float rounding of the subtraction only)
    assert float((target.double() * sigma * sigma + (noisy - f[:, None, :]).double()).abs().max()) < 1e-5
    # deterministic under the seed + site; a second call draws new numbers
    noisy_b, _ = X.add_feature_noise(f, sigma=sigma, n_nodes=N)
    torch.manual_seed(11)
    XF._drop.site = 0
    noisy_c, _ = X.add_feature_noise(f, sigma=sigma, n_nodes=N)
    assert torch.equal(noisy, noisy_c) and not torch.equal(noisy, noisy_b)
    # gradient of the broadcast form: sum over the N copies
    fr = f.clone().requires_grad_(True)
    out, _ = X.add_feature_noise(fr, sigma=sigma, n_nodes=N)
    c = torch.randn_like(out)
    (out * c).sum().backward()
    _close(fr.grad, c.sum(1).cpu(), 1e-6, "broadcast grad")
    # full [B,N,H] form + planes hand-over
    full = torch.randn(4, N, H, device=dev())
    nz, _ = X.add_feature_noise(full, sigma=sigma)
    assert nz.shape == full.shape and XF._planes_of(nz) is not None
    assert abs(float(((nz - full) / sigma).std()) - 1.0) < 2e-2


@pytest.mark.parametrize("gnn,B,N", [("GCN", 64, 64), ("GCN", 40, 100), ("GIN", 48, 64), ("GIN", 24, 100)])
def test_wider_graphs_multi_tile_vs_fp64_oracle(gnn, B, N):
    """BASELINE configs[3] node counts at sizes that span many row tiles (N=64: unaligned message-passing tiles, two
    graphs per Gram tile; N=100: graph-aligned tiles, one graph per tile; M = 4096 / 4000 / 3072 / 2400 rows -> the
    CTA-pair projections, grouped launches and split-K weight gradients are all active), H=768, train mode with
    injected masks: generator outputs, input gradients and parameter gradients against the fp64 oracle."""
    import xggm_b200 as X
    from xggm_b200.functional import inject_keep_masks
    H = 768
    p = O.make_params(61 + N, gnn, H, 2, N, heads=False)
    cls = {"GCN": X.GCNGenerator, "GIN": X.GINGenerator}[gnn]
    mod = _load_params(cls(H, 2), p, "generator.").to(dev()).train()
    visn, _, adj_true = O.make_inputs(62 + N, B, N, H)
    adj = O.strip_diag(adj_true)
    nh = {"GCN": 3, "GIN": 2}[gnn]
    keeps = O.make_keeps(63, 2, nh, (B, N, H))
    x = visn.clone().to(dev()).requires_grad_(True)
    a = adj.clone().to(dev()).requires_grad_(True)
    with inject_keep_masks([m for layer in keeps for m in layer]):
        xo, ao = mod(x, a)
    g = torch.Generator().manual_seed(64)
    cx, ca = torch.randn(B, N, H, generator=g), torch.randn(B, N, N, generator=g)
    ((xo * cx.to(dev())).sum() + (ao * ca.to(dev())).sum()).backward()
    p64 = {k: v.double().requires_grad_(True) for k, v in p.items()}
    x64, a64 = visn.double().requires_grad_(True), adj.double().requires_grad_(True)
    fn = O.gcn_generator if gnn == "GCN" else O.gin_generator
    xr, ar = fn(x64, a64, p64, 2, keeps, pre="generator.")
    ((xr * cx.double()).sum() + (ar * ca.double()).sum()).backward()
    _close(xo, xr.detach(), TOL, "x_out")
    _close(ao, ar.detach(), TOL, "adj_out")
    assert float(torch.diagonal(ao, dim1=1, dim2=2).abs().max()) == 0.0
    _close(x.grad, x64.grad, 2 * TOL, "gx")
    _close(a.grad, a64.grad, 2 * TOL, "gadj")
    for k, v in mod.named_parameters():
        ref = p64["generator." + k].grad
        if v.numel() == 1:       # GIN eps: cancellation-heavy scalar (see SCALAR_TOL_TC)
            assert abs(float(v.grad) - float(ref)) <= SCALAR_TOL_TC * max(abs(float(ref)), 1.0), k
        else:
            _close(v.grad, ref, 3 * TOL, "g/" + k)
