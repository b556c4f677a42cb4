"""The stock-torch LXMERT stand-in of the iteration benchmark (tools/lxmert_torch.py, harness only) has the
reference's architecture: same state_dict keys / shapes / parameter count, and -- when the reference modules are
vendored in oracle/_ref -- the same outputs from the reference class's own weights."""
import os
import sys

import pytest
import torch

from conftest import ROOT, rel_l2

sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_parameter_count_matches_survey():
    import lxmert_torch as LT
    m = LT.LXRTFeatureExtractionTorch(LT.Config(vocab_size=30522))
    n = sum(p.numel() for p in m.parameters())
    assert abs(n - 207.9e6) < 0.1e6, n          # 207.9 M measured on the reference class


def test_outputs_match_reference_class():
    from oracle.ref_step import reference_available, _import_reference
    if not reference_available():
        pytest.skip("oracle/_ref not vendored")
    _import_reference()
    import lxrt.modeling as RM
    import lxmert_torch as LT
    old = (RM.VISUAL_CONFIG.l_layers, RM.VISUAL_CONFIG.x_layers, RM.VISUAL_CONFIG.r_layers)
    RM.VISUAL_CONFIG.l_layers, RM.VISUAL_CONFIG.x_layers, RM.VISUAL_CONFIG.r_layers = 2, 2, 1
    try:
        torch.manual_seed(0)
        cfg = RM.BertConfig(vocab_size_or_config_json_file=500, hidden_size=64, num_hidden_layers=2,
                            num_attention_heads=4, intermediate_size=128)
        ref = RM.LXRTFeatureExtraction(cfg, mode="lxr").eval()
    finally:
        RM.VISUAL_CONFIG.l_layers, RM.VISUAL_CONFIG.x_layers, RM.VISUAL_CONFIG.r_layers = old
    mine = LT.LXRTFeatureExtractionTorch(LT.Config(vocab_size=500, hidden_size=64, num_attention_heads=4,
                                                   intermediate_size=128, l_layers=2, x_layers=2, r_layers=1)).eval()
    mine.load_state_dict(ref.state_dict(), strict=True)            # identical keys and shapes
    B = 3
    g = torch.Generator().manual_seed(2)
    ids = torch.randint(1, 500, (B, 20), generator=g)
    mask = torch.ones(B, 20, dtype=torch.long)
    mask[0, 12:] = 0
    mask[2, 5:] = 0
    ids = ids * mask
    feats, boxes = torch.relu(torch.randn(B, 36, 2048, generator=g)), torch.rand(B, 36, 4, generator=g)
    with torch.no_grad():
        (l_r, v_r), p_r = ref(ids, None, mask, visual_feats=(feats, boxes))
        (l_m, v_m), p_m = mine(ids, None, mask, visual_feats=(feats, boxes))
    assert rel_l2(l_m, l_r) < 1e-5 and rel_l2(v_m, v_r) < 1e-5 and rel_l2(p_m, p_r) < 1e-5
