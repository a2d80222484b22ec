"""Drop-in counterparts of the reference's ``utils/layers.py`` operators, on CUDA tensors.

``attn_head`` (utils/layers.py:7-46) and ``SimpleAttLayer`` (utils/layers.py:132-164) keep the
reference's names, positional order and argument meaning.  Differences forced by the move from a
TF1 graph to eager PyTorch are keyword-only: ``params=`` (the variables TF would create
implicitly) and, for ``bias_mat``, that a ``MetaPathGraph`` (from ``process.adj_to_bias``) is
accepted next to the dense (1,N,N) reference bias.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib, ops
from .graph import MetaPathGraph


def elu(x):
    """Marker for tf.nn.elu (the kernel applies it in its epilogue)."""
    return torch.nn.functional.elu(x)


def identity(x):
    return x


class EdgeCoefs:
    """Attention coefficients restricted to edges: what ``return_coef=True`` yields
    (utils/layers.py:43-44) without the N x N zeros.  ``to_dense()`` gives the reference layout."""

    def __init__(self, graph: MetaPathGraph, alpha: torch.Tensor):
        self.graph = graph
        self.alpha = alpha  # (nnz, K)

    def to_dense(self, head: int = 0) -> torch.Tensor:
        n, m = self.graph.n_rows, self.graph.n_cols
        deg = (self.graph.indptr[1:] - self.graph.indptr[:-1])
        rows = torch.repeat_interleave(torch.arange(n, device=self.alpha.device), deg)
        out = torch.zeros(n, m, dtype=self.alpha.dtype, device=self.alpha.device)
        out[rows, self.graph.indices.long()] = self.alpha[:, head]
        return out.unsqueeze(0)


def as_graph(bias_mat, device=None) -> MetaPathGraph:
    """Accepts what reference call sites pass as ``bias_mat``: a MetaPathGraph, or a dense
    (1,N,N)/(N,N) bias (numpy or torch) with 0 on edges and -1e9 elsewhere."""
    if isinstance(bias_mat, MetaPathGraph):
        return bias_mat
    return MetaPathGraph.from_dense_bias(bias_mat, device=device)


def _squeeze_batch(seq: torch.Tensor) -> torch.Tensor:
    if seq.dim() == 3:
        if seq.shape[0] != 1:
            raise ValueError("batch_size must be 1 (the whole graph), as in ex_acm3025.py:21")
        return seq[0]
    return seq


_DROP_SEEDS = {}


def _drop_seed(device) -> torch.Tensor:
    """Device-resident dropout seed word for stand-alone attn_head calls (one per device)."""
    key = str(device)
    if key not in _DROP_SEEDS:
        _DROP_SEEDS[key] = torch.randint(0, 2 ** 31 - 1, (1,), dtype=torch.int32).to(device)
    return _DROP_SEEDS[key]


def new_head_params(F: int, H: int, device, generator=None) -> Dict[str, torch.nn.Parameter]:
    """The variables one ``attn_head`` call creates (layers.py:20,23,24,35) with TF's default
    initialisers: glorot-uniform kernels, zero biases."""
    def glorot(shape, fi, fo):
        lim = math.sqrt(6.0 / (fi + fo))
        return torch.empty(shape).uniform_(-lim, lim, generator=generator).to(device)
    mk = torch.nn.Parameter
    return {"W": mk(glorot((F, H), F, H)), "a1": mk(glorot((H,), H, 1)), "b1": mk(torch.zeros((), device=device)),
            "a2": mk(glorot((H,), H, 1)), "b2": mk(torch.zeros((), device=device)),
            "bias": mk(torch.zeros(H, device=device))}


def attn_head(seq, out_sz, bias_mat, activation, in_drop=0.0, coef_drop=0.0, residual=False,
              return_coef=False, *, params: Optional[Dict[str, torch.Tensor]] = None):
    """One attention head (utils/layers.py:7-46): ``activation(softmax(leaky_relu(f1 + f2^T) +
    bias_mat) @ (seq W) + bias)``, evaluated over the edges of ``bias_mat`` only.

    seq (1,N,F) fp32 CUDA; returns (1,N,out_sz) [and EdgeCoefs when ``return_coef``].
    ``params``: dict W (F,H), a1 (H,), b1 (), a2 (H,), b2 (), bias (H,); created with the
    reference initialisers when omitted and returned as ``attn_head.last_params``.
    """
    x = _squeeze_batch(seq)
    _lib.require_cuda(x)
    graph = as_graph(bias_mat, x.device)
    H = int(out_sz)
    if params is None:
        params = new_head_params(x.shape[1], H, x.device)
    attn_head.last_params = params
    use_res = bool(residual) and x.shape[1] != H                     # :39-40; equal widths: a dead store (:42)
    if use_res:
        if "W_res" not in params:                                    # conv1d(seq, H, 1): glorot kernel, zero bias
            lim = math.sqrt(6.0 / (x.shape[1] + H))
            params["W_res"] = torch.nn.Parameter(torch.empty(x.shape[1], H).uniform_(-lim, lim).to(x.device))
            params["b_res"] = torch.nn.Parameter(torch.zeros(H, device=x.device))
    seed = None
    if in_drop or coef_drop:
        seed = ops.next_seed(_drop_seed(x.device))                   # this call's own snapshot (fwd + bwd masks)
    act = ops.activation_code(activation)
    plan = ops.NodeAttentionPlan(graphs=[graph], K=1, H=H, act=act, want_coefs=bool(return_coef),
                                 in_drop=float(in_drop), coef_drop=float(coef_drop), seed=seed)
    res, bias = None, params["bias"]
    if use_res:
        # ret + conv1d(seq, H, 1) before the activation (:40); `seq` there is the DROPPED input of :19, i.e. the
        # same mask stream as this head's projection.  The conv's bias is folded into the head bias.
        res = ops.residual_conv(x, params["W_res"], 1, H, seed, float(in_drop), 0, graph.row_offset)
        bias = bias + params["b_res"]
    Z = ops.node_attention(plan, x, params["W"], params["a1"].reshape(1, 1, H), params["b1"].reshape(1, 1),
                           params["a2"].reshape(1, 1, H), params["b2"].reshape(1, 1), bias.reshape(1, H), res)
    ret = Z.reshape(1, x.shape[0], H)
    if return_coef:
        return ret, EdgeCoefs(graph, plan.coefs[0])
    return ret


attn_head.last_params = None


def attn_head_const_1(seq, out_sz, bias_mat, activation, in_drop=0.0, coef_drop=0.0, residual=False, *,
                      params: Optional[Dict[str, torch.Tensor]] = None):
    """The HAN_nd ablation head (utils/layers.py:49-81): logits = the 0/1 adjacency itself, so after the
    -1e9 mask every neighbour gets the same weight 1/deg(i) -- node-level attention switched off.
    Runs the same kernels with constant scores (a1 = a2 = 0, b1 = 0, b2 = 1: e_ij = leaky_relu(1) on every
    edge); only W and the output bias are variables (:56,70)."""
    x = _squeeze_batch(seq)
    _lib.require_cuda(x)
    H = int(out_sz)
    if params is None:
        full = new_head_params(x.shape[1], H, x.device)
        params = {"W": full["W"], "bias": full["bias"]}
    attn_head_const_1.last_params = params
    dev = x.device
    const = {"W": params["W"], "bias": params["bias"], "a1": torch.zeros(H, device=dev), "b1": torch.zeros((), device=dev),
             "a2": torch.zeros(H, device=dev), "b2": torch.ones((), device=dev)}
    for k in ("W_res", "b_res"):
        if k in params:
            const[k] = params[k]
    out = attn_head(seq, out_sz, bias_mat, activation, in_drop=in_drop, coef_drop=coef_drop, residual=residual,
                    params=const)
    for k in ("W_res", "b_res"):
        if k in const:
            params[k] = const[k]
    return out


attn_head_const_1.last_params = None


def sp_attn_head(seq, out_sz, adj_mat, activation, nb_nodes, in_drop=0.0, coef_drop=0.0, residual=False, *,
                 params: Optional[Dict[str, torch.Tensor]] = None):
    """Sparse-adjacency head (utils/layers.py:85-127): logits ``adj_ij * f1_i + adj_ij * f2_j`` on the STORED
    entries (:95-96), leaky_relu, ``tf.sparse_softmax`` over each row's stored entries (:100), sparse @ dense
    (:113).  The stored values w_ij ride along as one extra 4-byte stream per edge through K-B / K-D; with the
    0/1 adjacency the reference's drivers build, this is ``attn_head`` exactly.

    adj_mat: a torch sparse COO/CSR tensor (N,N) or (1,N,N) -- its stored values are the weights -- or a
    ``MetaPathGraph`` (optionally carrying ``edge_weight``)."""
    x = _squeeze_batch(seq)
    if isinstance(adj_mat, torch.Tensor) and adj_mat.layout != torch.strided:
        coo = adj_mat.to_sparse_coo().coalesce()
        idx, vals = coo.indices(), coo.values()
        if idx.shape[0] == 3:                                        # batch-size-1 SparseTensor, :110-113
            if int(coo.shape[0]) != 1:
                raise ValueError("sp_attn_head assumes batch size 1 (utils/layers.py:110-113)")
            idx = idx[1:]
        n = int(nb_nodes)
        order = torch.argsort(idx[0] * n + idx[1])
        rows, cols = idx[0][order], idx[1][order]
        indptr = torch.zeros(n + 1, dtype=torch.int64, device=rows.device)
        indptr[1:] = torch.cumsum(torch.bincount(rows, minlength=n), 0)
        graph = MetaPathGraph.from_csr(indptr.to(x.device), cols.to(torch.int32).to(x.device), n_cols=n, device=x.device)
        w = vals[order].to(torch.float32).to(x.device).contiguous()
        if w.numel() and not bool((w == 1).all()):
            graph.edge_weight = w
        adj_mat = graph
    elif not isinstance(adj_mat, MetaPathGraph):
        raise TypeError("adj_mat: MetaPathGraph or a torch sparse tensor")
    if adj_mat.has_empty_rows():
        raise ValueError("sp_attn_head: a row without stored entries has no softmax (the reference's adjacency "
                         "carries self-loops)")
    return attn_head(seq, out_sz, adj_mat, activation, in_drop=in_drop, coef_drop=coef_drop, residual=residual,
                     params=params)


def SimpleAttLayer(inputs, attention_size, time_major=False, return_alphas=False, *,
                   params: Optional[Dict[str, torch.Tensor]] = None, mode: str = "reference", dist=None):
    """Semantic-level attention (utils/layers.py:132-164).  inputs (N,P,D) -> (N,D) [+ alphas (N,P)].

    ``mode="reference"`` is the shipped per-node softmax over meta-paths (:156);
    ``mode="paper"`` averages the scores over all nodes first (han.pdf Eq. 7-9).
    """
    if isinstance(inputs, tuple):                       # :134-136
        inputs = torch.cat(inputs, 2)
    if time_major:                                      # :138-140
        inputs = inputs.transpose(0, 1)
    _lib.require_cuda(inputs)
    D = inputs.shape[2]
    if params is None:
        dev = inputs.device
        params = {"w_omega": torch.nn.Parameter((torch.randn(D, attention_size) * 0.1).to(dev)),   # :145
                  "b_omega": torch.nn.Parameter((torch.randn(attention_size) * 0.1).to(dev)),      # :146
                  "u_omega": torch.nn.Parameter((torch.randn(attention_size) * 0.1).to(dev))}      # :147
    SimpleAttLayer.last_params = params
    assert params["w_omega"].shape == (D, attention_size)
    out, alphas = ops.semantic_attention(inputs, params["w_omega"], params["b_omega"], params["u_omega"],
                                         mode=mode, dist=dist)
    if not return_alphas:
        return out
    return out, alphas


SimpleAttLayer.last_params = None
