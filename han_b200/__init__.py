"""han_b200 — B200-native (sm_100a) implementation of HAN's node-level and semantic-level attention
behind the reference's operator surface (CG-Labs/HAN: utils/process.adj_to_bias, utils/layers.attn_head,
utils/layers.SimpleAttLayer, models/gat.HeteGAT_multi.inference).

CUDA-only: every op calls hand-written kernels in libhan_sm100.so through ctypes
(include/han_b200.h); there is no CPU or PyTorch-math fallback.
"""
from . import _lib, base_gattn, gat, graph, layers, ops, process, variables  # noqa: F401
from .base_gattn import BaseGAttN  # noqa: F401
from .gat import GAT, HeteGAT, HeteGAT_multi  # noqa: F401
from .graph import MetaPathGraph  # noqa: F401
from .variables import GATParams, HANParams  # noqa: F401

__all__ = ["GAT", "GATParams", "HeteGAT", "HeteGAT_multi", "BaseGAttN", "MetaPathGraph", "HANParams", "layers", "process"]
