"""xggm_b200 -- B200-native (sm_100a) X-GGM graph generative block.

Drop-in for the reference's ``module.graph_generative_modeling`` / ``module.gcn`` /
``module.gin`` / ``module.gat`` / ``module.graph_utils`` hot path; see DESIGN.md.
Importing the package does not need a GPU; running any operator does (no fallback).
"""
from . import _lib, data, ddp, functional, glue, graphs, optim  # noqa: F401
from .glue import (add_edge_noise, add_edge_noise_v2, add_feature_noise,  # noqa: F401
                   add_feature_noise_v2, compute_kl_loss, loss_func)
from ._lib import get_precision, set_precision  # noqa: F401
from .graphs import GraphedStep  # noqa: F401
from .model import AnswerHead, XGGMHeads  # noqa: F401
from .optim import BertAdam, bce_with_logits, clip_grad_norm_  # noqa: F401
from .nn import (GAT, GCN, GIN, Discriminator, DiscriminatorV2, EdgeGenerator, GATConv,  # noqa: F401
                 GATGenerator, GCNConv, GCNGenerator, GCNPlainEncoder, GeLU, GINConv, GINGenerator,
                 GinPlainEncoder, MixGenerator, NodeGenerator, VisualFeatEncoder)

__version__ = "0.1.0"
