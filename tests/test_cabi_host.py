"""CPU-side checks: the C-ABI library builds, loads and exports every symbol the header
declares (no compute calls without a GPU); host logic of the boundary."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT
from oracle import xggm_oracle as O


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "xggm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(xggm_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_loads_and_exports_header_symbols():
    import __graft_entry__ as G
    G.build()
    from xggm_b200 import _lib
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in xggm_b200.h but not exported"
    # every declared entry point has a ctypes signature (argument-count drift is a bug)
    assert set(syms) == set(_lib.SIGNATURES), set(syms) ^ set(_lib.SIGNATURES)
    assert lib.xggm_abi_version() == 4
    assert b"argument" in lib.xggm_strerror(-1)
    # 2 convs x (3 fp32 [M,H] + 2 [M] + 2 plane regions) + 3 heads x (z + mean + rstd) + P(x); M = 72
    assert lib.xggm_gnn_saved_floats(0, 2, 36, 768, 2) == 2 * (5 * 55296 + 144) + 3 * (55296 + 144) + 55296
    assert lib.xggm_linear_work_bytes(72, 768, 768) == 4 * (2 * 55296 + 768 * 768)
    assert lib.xggm_gnn_saved_floats(7, 2, 36, 768, 2) == -1


def test_signature_arity_matches_header():
    from xggm_b200 import _lib
    src = open(os.path.join(ROOT, "include", "xggm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for name, args in re.findall(r"\b(xggm_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src):
        n = 0 if args.strip() in ("", "void") else len(args.split(","))
        assert len(_lib.SIGNATURES[name]) == n, name


def test_only_sm100a_sass_in_library():
    from xggm_b200 import _lib
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback():
    import xggm_b200 as X
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(RuntimeError, match="no CPU path"):
        X.GCNGenerator(64, 1)(torch.zeros(1, 36, 64), torch.zeros(1, 36, 36))
    with pytest.raises(RuntimeError, match="no CPU path"):
        X.loss_func(torch.zeros(1, 4, 4), torch.zeros(1, 4, 4), 1.0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "xggm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+\.*oracle", txt, flags=re.M), f
                assert "import_module(\"oracle" not in txt and "/root/reference" not in txt, f


@pytest.mark.parametrize("gnn,L", [("GCN", 2), ("GIN", 2), ("GAT", 2), ("GCN", 1)])
def test_state_dict_layout_matches_reference(gnn, L):
    """Names/shapes probed from the reference modules (oracle.param_shapes is pinned by the
    fixtures, whose parameters were loaded into the reference with strict=True)."""
    import xggm_b200 as X
    m = X.XGGMHeads(768, gnn, L)
    got = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    assert got == O.param_shapes(gnn, 768, L, 36, heads=True)
    n_gen = sum(v.numel() for k, v in m.state_dict().items() if k.startswith("generator."))
    assert n_gen == {("GCN", 2): 5918208, ("GIN", 2): 3552770, ("GAT", 2): 2365440, ("GCN", 1): 2959104}[(gnn, L)]


def test_default_init_matches_reference_recipe():
    """nn.Linear kaiming-uniform / LayerNorm ones-zeros (the reference's reset_parameters is
    commented out, src/module/gcn.py:17); GATConv xavier_normal gain sqrt(2) (gat.py:20-23)."""
    import xggm_b200 as X
    torch.manual_seed(0)
    g = X.GCNGenerator(768, 1)
    w = g.gnn_layers[0].gnn_layers[0].ctx_layer.weight
    assert float(w.abs().max()) <= 1 / 768 ** 0.5 + 1e-6
    ln = g.gnn_layers[0].gnn_layers[0].layer_norm
    assert torch.equal(ln.weight, torch.ones(768)) and torch.equal(ln.bias, torch.zeros(768)) and ln.eps == 1e-5
    a = X.GATConv(768, 768)
    std = float(a.linear_layer.weight.std())
    assert abs(std - (2.0 ** 0.5) * (2.0 / (768 + 768)) ** 0.5) < 2e-3
    assert float(X.GINConv(8, 8).eps) == 0.0


def test_unequal_widths_rejected():
    import xggm_b200 as X
    with pytest.raises(NotImplementedError):
        X.GCN(768, [512, 768], 2)


def test_mask_injection_bookkeeping():
    from xggm_b200.functional import inject_keep_masks, _mask_feed
    with pytest.raises(RuntimeError, match="not consumed"):
        with inject_keep_masks([torch.ones(2, 2, dtype=torch.uint8)]):
            pass
    assert _mask_feed == []


def test_bertadam_host_logic_flat_layout_and_schedule():
    """xggm_b200.optim.BertAdam on CPU tensors: parameters and gradients share one offset table, argument
    validation mirrors the reference constructor, the schedule is the reference's (no kernels run here)."""
    import pytest
    import torch
    from oracle import xggm_oracle as O
    from xggm_b200.ddp import FlatGrads
    from xggm_b200.optim import BertAdam, SCHEDULES
    ps = [torch.nn.Parameter(torch.randn(5, 7)), torch.nn.Parameter(torch.randn(3)), torch.nn.Parameter(torch.randn(2, 2))]
    before = [p.detach().clone() for p in ps]
    fg = FlatGrads(ps)
    opt = BertAdam(ps, lr=1e-3, warmup=0.1, t_total=50, flat_grads=fg)
    grp = opt.groups[0]
    assert grp.grads is fg and grp.flat_p.shape == fg.flat.shape
    for p, b in zip(ps, before):
        assert torch.equal(p.detach(), b)                                   # values survive the move
        off_p = (p.data_ptr() - grp.flat_p.data_ptr()) // 4
        off_g = (p.grad.data_ptr() - fg.flat.data_ptr()) // 4
        assert off_p == off_g and off_p % 32 == 0                           # same 128-byte aligned offsets
    assert opt.get_lr() == [0]
    for step in (0, 3, 5, 20, 49, 60):
        grp.step_dev.fill_(step)
        assert grp.step == step
        assert grp.lr_scheduled() == O.scheduled_lr(1e-3, step, 50, 0.1)
    grp.step_dev.fill_(0)
    for name, fn in SCHEDULES.items():
        for x in (0.0, 0.001, 0.3, 1.0, 1.2):
            assert fn(x, 0.25) == O.SCHEDULES[name](x, 0.25)
    with pytest.raises(ValueError):
        BertAdam(ps, lr=-1.0)
    with pytest.raises(ValueError):
        BertAdam(ps, lr=1e-3, schedule="nope")
    with pytest.raises(ValueError):
        BertAdam(ps, lr=1e-3, b1=1.0)
    with pytest.raises(ValueError):
        BertAdam([torch.nn.Parameter(torch.zeros(2))], lr=1e-3, flat_grads=fg)   # bucket of other parameters


def test_operand_plane_record_is_dropped_when_the_tensor_changes():
    """functional._attach_planes / _planes_of (host logic of the operand-plane hand-over): the record follows
    the tensor object and is invalidated by an in-place update, a re-allocation or an engine change."""
    import torch
    import xggm_b200 as X
    import xggm_b200.functional as XF
    t = torch.zeros(4, 6, 16)
    planes = torch.empty(64, dtype=torch.uint8)
    assert XF._planes_of(t) is None
    XF._attach_planes(t, planes)
    assert XF._planes_of(t) is planes
    assert XF._planes_of(t.clone()) is None            # a different tensor object / storage
    X.set_precision("bf16")                              # planes written under another engine are not reused
    try:
        assert XF._planes_of(t) is None
    finally:
        X.set_precision("fp32")
    assert XF._planes_of(t) is planes
    t.add_(1.0)                                          # in-place update bumps the version counter
    assert XF._planes_of(t) is None
    assert XF._new_planes(torch.zeros(2, 3, 10)) is None     # row length not a multiple of 8: producers emit nothing


def test_flatgrads_survives_model_zero_grad_and_tracks_active_parameters():
    """model.zero_grad() (set_to_none on torch >= 2) un-links every p.grad from the bucket; the next
    backward then allocates gradients elsewhere.  relink() -- run by all_reduce() and BertAdam.step() -- copies
    them into their slices and restores the link; strict mode raises.  active_ranges() lists the bucket ranges of
    the parameters that received a gradient since zero_() (the reference optimiser skips p.grad is None)."""
    import torch
    from xggm_b200.ddp import FlatGrads
    model = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2), torch.nn.Linear(2, 5))
    fg = FlatGrads(model.parameters())
    x = torch.randn(6, 4)
    model[1](model[0](x)).sum().backward()                # the third layer takes no part
    assert float(fg.flat.abs().sum()) > 0
    used = sum(o2 - o1 for (o1, o2) in fg.active_ranges())
    assert fg.active_ranges()[0][0] == 0 and used == fg.offsets[4]          # four tensors of the first two layers
    ref = fg.flat.clone()
    model.zero_grad()                                      # what the reference trainer calls (vqacpv2.py:170)
    assert all(p.grad is None for p in model.parameters())
    fg.zero_()
    model[1](model[0](x)).sum().backward()
    assert float(fg.flat.abs().sum()) == 0.0               # the gradients live outside the bucket now ...
    with pytest.raises(RuntimeError, match="no longer points"):
        fg.relink(strict=True)
    assert fg.relink() == 6                                # ... until the link is repaired (all six were un-linked)
    assert torch.equal(fg.flat, ref)
    base = fg.flat.data_ptr()
    for p, o in zip(fg.params, fg.offsets):
        assert p.grad.data_ptr() == base + 4 * o
    assert sum(o2 - o1 for (o1, o2) in fg.active_ranges()) == used          # the unused layer stays inactive
    assert fg.relink() == 0
    fg.zero_()
    assert fg.active_ranges() == [(0, fg.flat.numel())]    # nothing reported: everything counts as active


def test_fused_grad_accumulation_is_opt_in():
    """only parameters registered with a FlatGrads may be accumulated into in place."""
    import torch
    import xggm_b200.functional as XF
    from xggm_b200.ddp import FlatGrads
    p = torch.nn.Parameter(torch.randn(3, 3))
    p.grad = torch.zeros(3, 3)
    assert XF._grad_target(p) is None                      # a dense .grad alone is not consent
    fg = FlatGrads([p])
    assert XF._grad_target(p) is p.grad
    p.grad = torch.zeros(3, 3)                             # un-linked from the bucket
    assert XF._grad_target(p) is None
    fg.relink()
    XF.FUSE_GRAD_ACCUMULATION = False
    try:
        assert XF._grad_target(p) is None
    finally:
        XF.FUSE_GRAD_ACCUMULATION = True


def test_bertadam_state_dict_round_trip():
    import torch
    from xggm_b200.optim import BertAdam
    ps = [torch.nn.Parameter(torch.randn(5, 7)), torch.nn.Parameter(torch.randn(3))]
    opt = BertAdam(ps, lr=1e-3, warmup=0.1, t_total=50)
    g = opt.groups[0]
    g.m.normal_(); g.v.uniform_(); g.step_dev.fill_(7)
    sd = opt.state_dict()
    assert set(sd) == {"state", "param_groups"} and sd["state"][0]["step"] == 7
    assert sd["state"][0]["next_m"].shape == (5, 7) and sd["param_groups"][0]["params"] == [0, 1]
    ps2 = [torch.nn.Parameter(torch.randn(5, 7)), torch.nn.Parameter(torch.randn(3))]
    opt2 = BertAdam(ps2, lr=5e-4)
    opt2.load_state_dict(sd)
    g2 = opt2.groups[0]
    for p, o in zip(g2.params, g2.grads.offsets):
        n = p.numel()
        assert torch.equal(g2.m[o:o + n], g.m[o:o + n]) and torch.equal(g2.v[o:o + n], g.v[o:o + n])
    assert g2.step == 7 and g2.opts["lr"] == 1e-3 and g2.opts["t_total"] == 50


def test_bertadam_groups_sharing_one_bucket_keep_their_learning_rates():
    """Two parameter groups (the trainers' encoder / down-task split, src/vqa/vqacpv2.py:113-128) in ONE FlatGrads:
    one internal group with a per-parameter base lr, update ranges split where the lr changes and restricted to the
    parameters that received a gradient."""
    import torch
    from xggm_b200.ddp import FlatGrads
    from xggm_b200.optim import BertAdam
    enc = [torch.nn.Parameter(torch.randn(40, 8)), torch.nn.Parameter(torch.randn(8))]
    down = [torch.nn.Parameter(torch.randn(16, 8)), torch.nn.Parameter(torch.randn(3)), torch.nn.Parameter(torch.randn(5))]
    fg = FlatGrads(enc + down)
    opt = BertAdam([{"params": enc, "lr": 1e-3}, {"params": down, "lr": 4e-3}], lr=1e-3, warmup=0.1, t_total=10, flat_grads=fg)
    assert len(opt.groups) == 1 and [g["lr"] for g in opt.param_groups] == [1e-3, 4e-3]
    assert [len(g["params"]) for g in opt.param_groups] == [2, 3]
    g = opt.groups[0]
    r = g.active_ranges()                      # nothing reported: everything active, split at the lr boundary
    assert len(r) == 2 and r[0][0] == 0 and r[0][1] == r[1][0] == fg.offsets[2] and r[1][1] == fg.flat.numel()
    assert (r[0][2], r[1][2]) == (1e-3, 4e-3)
    fg.zero_()
    (enc[0].sum() + down[1].sum() + down[2].sum()).backward()     # enc[1] and down[0] take no part
    r = g.active_ranges()
    assert [(a, b) for a, b, _ in r] == [(fg.offsets[0], fg.offsets[1]), (fg.offsets[3], fg.flat.numel())]
    assert [lr for _, _, lr in r] == [1e-3, 4e-3]
    g.step_dev.fill_(3)
    lrs = opt.get_lr()
    assert len(lrs) == 5 and abs(lrs[2] - 4 * lrs[0]) < 1e-12
    with pytest.raises(ValueError, match="lr only"):
        BertAdam([{"params": enc}, {"params": down, "weight_decay": 0.5}], lr=1e-3, flat_grads=FlatGrads(enc + down))
