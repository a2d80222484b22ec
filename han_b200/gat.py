"""Drop-in counterpart of the reference's ``models/gat.py`` ``HeteGAT_multi`` (models/gat.py:34-77)."""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import _lib, layers, ops, variables
from .base_gattn import BaseGAttN
from .graph import MetaPathGraph
from .layers import elu


def _group_by_input(xs: Sequence[torch.Tensor]) -> List[List[int]]:
    """Meta-paths whose feature tensors are the same storage share one projection launch
    (ACM feeds the same features for every meta-path, ex_acm3025.py:86)."""
    groups, seen = [], {}
    for p, x in enumerate(xs):
        key = (x.data_ptr(), tuple(x.shape), tuple(x.stride()))
        if key in seen:
            groups[seen[key]].append(p)
        else:
            seen[key] = len(groups)
            groups.append([p])
    return groups


class HeteGAT(BaseGAttN):
    """models/gat.py:132-203: one shared feature tensor for every meta-path and, with
    ``return_coef=True``, the head-averaged attention coefficients of every meta-path
    (``tf.concat(head_coef_list, 0)`` -> ``reduce_mean(axis=0)``, :165-167) -- the interpretability
    output of the paper.  The coefficients come back restricted to edges (``layers.EdgeCoefs`` with
    one column; ``.to_dense()`` gives the reference's (1,N,N) matrix)."""

    @staticmethod
    def inference(inputs, nb_classes, nb_nodes, training, attn_drop, ffd_drop,
                  bias_mat_list, hid_units, n_heads, activation=None, residual=False,
                  mp_att_size=128, return_coef=False, **kw):
        out = HeteGAT_multi.inference([inputs] * len(bias_mat_list), nb_classes, nb_nodes, training, attn_drop,
                                      ffd_drop, bias_mat_list, hid_units, n_heads,
                                      activation=elu if activation is None else activation, residual=residual,
                                      mp_att_size=mp_att_size, return_coef=return_coef, **kw)
        if not return_coef:
            return out                                                  # :202-203
        logits, final_embed, att_val, coefs = out
        coef_list = [layers.EdgeCoefs(c.graph, c.alpha.mean(dim=1, keepdim=True)) for c in coefs]   # :165-167
        return logits, final_embed, att_val, coef_list                 # :200-201


class HeteGAT_multi(BaseGAttN):
    @staticmethod
    def inference(inputs_list, nb_classes, nb_nodes, training, attn_drop, ffd_drop,
                  bias_mat_list, hid_units, n_heads, activation=elu, residual=False,
                  mp_att_size=128, *, params: Optional[variables.HANParams] = None,
                  semantic_mode: str = "reference", dist=None, project_mode: int = 0,
                  return_coef: bool = False):
        """Same positional signature as models/gat.py:35-37; returns
        ``(logits (1,N,C), final_embed (N,D), att_val (N,P))`` (:76-77).

        inputs_list: P feature tensors (1,N,F) fp32 CUDA; bias_mat_list: P ``MetaPathGraph``s (from
        ``process.adj_to_bias``) or dense reference biases.  ``zip`` truncation of the two lists
        (:39) is kept.  ``nb_nodes`` and ``training`` are accepted and ignored, as in the reference.
        Keyword-only extras: ``params`` (else the process-wide default store, like TF's default
        graph), ``semantic_mode`` ("reference" per-node beta | "paper" node-mean beta), ``dist``
        (row-shard context for multi-GPU), ``project_mode``, ``return_coef``.
        """
        attn_drop, ffd_drop = float(attn_drop), float(ffd_drop)
        if not (0.0 <= attn_drop < 1.0 and 0.0 <= ffd_drop < 1.0):
            raise ValueError("attn_drop and ffd_drop are probabilities of dropping, in [0, 1)")
        if residual:
            raise NotImplementedError("residual=True is dead code for hid_units=[8] (gat.py:45) and not built")
        if len(hid_units) != 1:
            raise NotImplementedError("stacked attention layers (models/gat.py:48-57) are not built yet")
        pairs = list(zip(inputs_list, bias_mat_list))                      # :39
        P = len(pairs)
        xs = [layers._squeeze_batch(x) for x, _ in pairs]
        _lib.require_cuda(*xs)
        dev = xs[0].device
        graphs = [layers.as_graph(b, dev) for _, b in pairs]
        K, H = int(n_heads[0]), int(hid_units[0])
        if params is None:
            params = variables.get_default_store()
            if params is None:
                params = variables.HANParams([x.shape[1] for x in xs], nb_classes, hid_units, n_heads,
                                             mp_att_size, device=dev)
                variables.set_default_store(params)
        act = ops.activation_code(activation)
        seed = None
        if attn_drop or ffd_drop:
            # training-mode dropout (ex_acm3025.py:185-186 feeds 0.6 / 0.6; 0.0 at evaluation): a fresh mask
            # set per call; the seed word lives on the device so a captured CUDA graph advances it too
            seed = params.drop_seed
            seed.add_(1)

        coef_out = [None] * P
        groups = _group_by_input(xs)
        z_parts = []
        for grp in groups:
            plan = ops.NodeAttentionPlan(graphs=[graphs[p] for p in grp], K=K, H=H, act=act,
                                         project_mode=project_mode, dist=dist, want_coefs=return_coef,
                                         in_drop=ffd_drop, coef_drop=attn_drop, seed=seed, metapath_ids=list(grp))
            if len(grp) == 1:
                p = grp[0]
                W, a1, b1 = params.W[p], params.a1[p].unsqueeze(0), params.b1[p].unsqueeze(0)
                a2, b2, bias = params.a2[p].unsqueeze(0), params.b2[p].unsqueeze(0), params.bias[p].unsqueeze(0)
            else:
                W = torch.cat([params.W[p] for p in grp], dim=1)
                a1 = torch.stack([params.a1[p] for p in grp]); b1 = torch.stack([params.b1[p] for p in grp])
                a2 = torch.stack([params.a2[p] for p in grp]); b2 = torch.stack([params.b2[p] for p in grp])
                bias = torch.stack([params.bias[p] for p in grp])
            z_parts.append(ops.node_attention(plan, xs[grp[0]], W, a1, b1, a2, b2, bias))   # :42-58
            if return_coef:
                for p, alpha in zip(grp, plan.coefs):
                    coef_out[p] = layers.EdgeCoefs(graphs[p], alpha)
        if len(groups) == 1:
            multi_embed = z_parts[0]                                        # :60  (N,P,D)
        else:
            order = [p for grp in groups for p in grp]
            inv = sorted(range(P), key=lambda i: order[i])
            multi_embed = torch.cat(z_parts, dim=1)[:, inv, :].contiguous()

        final_embed, att_val = layers.SimpleAttLayer(                       # :61-63
            multi_embed, mp_att_size, time_major=False, return_alphas=True,
            params={"w_omega": params.w_omega, "b_omega": params.b_omega, "u_omega": params.u_omega},
            mode=semantic_mode, dist=dist)

        out = []
        for i in range(n_heads[-1]):                                        # :66-68
            out.append(torch.addmm(params.bc[i], final_embed, params.Wc[i]))
        logits = out[0] if len(out) == 1 else torch.stack(out).sum(0) / n_heads[-1]   # :72
        logits = logits.unsqueeze(0)                                        # :76
        if return_coef:
            return logits, final_embed, att_val, coef_out
        return logits, final_embed, att_val
