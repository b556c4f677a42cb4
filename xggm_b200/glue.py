"""Trainer-side glue of the GGM step, under the reference's own function names.

  add_edge_noise / add_feature_noise  = add_*_noise_v2, src/module/graph_utils.py:144-168
                                        (the aliases the trainers import, src/vqa/vqacpv2.py:18-19)
  loss_func / compute_kl_loss         = src/vqa/vqacpv2.py:48-61 (= src/gqa/gqa_ood.py:48-61)
  strip_diag / triu_scatter           = src/vqa/vqacpv2.py:188 / :195-199

Feature noise draws its Gaussians INSIDE the kernel (Philox4x32-10 + Box-Muller keyed by ``torch.initial_seed()``, a
per-call site number and -- under CUDA-graph replay -- the device epoch, like the library's dropout): no ``randn``
tensor, one launch less.  ``XGGM_TORCH_RANDN=1`` (or passing ``randn=``) draws with ``torch.randn`` instead, which
reproduces the reference's own generator stream on the same device.  Edge noise ([B,36,36]) keeps ``torch.randn_like``.
"""
import os

import torch

from . import functional as XF

strip_diag = XF.strip_diag
triu_scatter = XF.triu_scatter
fuse_readout = XF.fuse_readout


def add_edge_noise_v2(adjs, sigma=0.2, randn=None):
    assert isinstance(adjs, torch.Tensor)
    if randn is None:
        randn = torch.randn_like(adjs)
    return XF._EdgeNoise.apply(adjs, randn, sigma)


def add_feature_noise_v2(feats, sigma=0.2, randn=None, n_nodes=None):
    """feats [B,N,H]; or [B,H] with ``n_nodes`` to perturb N broadcast copies of each row
    (node_fc on 36 identical rows, src/vqa/vqacpv2.py:228-231) without materialising them."""
    assert isinstance(feats, torch.Tensor)
    if feats.dim() == 2:
        if n_nodes is None and randn is None:
            raise RuntimeError("xggm_b200: n_nodes is required for broadcast [B,H] features")
        shape = (feats.shape[0], n_nodes if randn is None else randn.shape[1], feats.shape[1])
    else:
        shape = feats.shape
    if randn is None and shape[-1] % 4 == 0 and os.environ.get("XGGM_TORCH_RANDN") != "1":
        planes = XF._new_planes(_shape_probe(shape, feats.device))
        noisy, target = XF.feat_noise_philox(feats, shape, sigma, planes)
        XF._attach_planes(noisy, planes)
        return noisy, target
    if randn is None:
        randn = torch.randn(shape, device=feats.device, dtype=feats.dtype)
    planes = XF._new_planes(randn)
    noisy, target = XF._FeatNoise.apply(feats, randn, sigma, planes)
    XF._attach_planes(noisy, planes)   # operand planes for the first GNN layer that consumes `noisy`
    return noisy, target


def _shape_probe(shape, device):
    """A zero-storage tensor of the given shape / device (``_new_planes`` only looks at shape, numel and device)."""
    return torch.empty(1, device=device).expand(shape)


add_edge_noise = add_edge_noise_v2
add_feature_noise = add_feature_noise_v2


def loss_func(score, grad_log_q_noise, sigma=0.2):
    """Score-matching loss: 0.5 sigma^2 mean_b sum_{last two dims}(score-target)^2 / (R*C)."""
    return XF._ScoreMse.apply(score, grad_log_q_noise, sigma)


def compute_kl_loss(x, y):
    """Symmetric KL between the last-dim softmaxes of x and y, mean over all elements."""
    return XF._SymKl.apply(x, y)
