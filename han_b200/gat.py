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


def _pad_heads(lay, K: int, H: int, Hp: int):
    """Zero-pad every head from H to Hp columns (the kernels are instantiated for H in {4, 8, 16}); the pad
    columns project to 0, add 0 to f1/f2 and come out as act(0) -- sliced away by the caller."""
    if Hp == H:
        return lay["W"], lay["a1"], lay["a2"], lay["bias"]
    pad = torch.nn.functional.pad
    F = lay["W"].shape[0]
    return (pad(lay["W"].view(F, K, H), (0, Hp - H)).reshape(F, K * Hp), pad(lay["a1"], (0, Hp - H)),
            pad(lay["a2"], (0, Hp - H)), pad(lay["bias"].view(K, H), (0, Hp - H)).reshape(K * Hp))


class GAT(BaseGAttN):
    """The homogeneous baseline, models/gat.py:8-32: ``len(hid_units)`` attention layers (heads concatenated),
    then ``n_heads[-1]`` output heads of width ``nb_classes`` with the identity activation, averaged (:25-30).
    Same kernels as HAN with one graph and no semantic layer."""

    @staticmethod
    def inference(inputs, nb_classes, nb_nodes, training, attn_drop, ffd_drop, bias_mat, hid_units, n_heads,
                  activation=elu, residual=False, *, params: Optional[variables.GATParams] = None,
                  project_mode: int = 0):
        attn_drop, ffd_drop = float(attn_drop), float(ffd_drop)
        x = layers._squeeze_batch(inputs)
        _lib.require_cuda(x)
        graph = layers.as_graph(bias_mat, x.device)
        if params is None:
            params = variables.GATParams(x.shape[1], nb_classes, hid_units, n_heads, device=x.device, residual=residual)
        seed = None
        if attn_drop or ffd_drop:
            seed = ops.next_seed(params.drop_seed)                  # this call's own snapshot (fwd + bwd masks)
        act = ops.activation_code(activation)

        def layer(h, lay, K, H, act_code, stream_id, mode):
            Hp = next(c for c in (4, 8, 16) if c >= H) if H <= 16 else H
            W, a1, a2, bias = _pad_heads(lay, K, H, Hp)
            res = None
            if "W_res" in lay:      # ret + conv1d(seq, H, 1) before the activation (layers.py:38-40), per-head dropped input
                pad = torch.nn.functional.pad
                Fin = lay["W_res"].shape[0]
                W_res = lay["W_res"] if Hp == H else pad(lay["W_res"].view(Fin, K, H), (0, Hp - H)).reshape(Fin, K * Hp)
                b_res = lay["b_res"] if Hp == H else pad(lay["b_res"].view(K, H), (0, Hp - H)).reshape(K * Hp)
                res = ops.residual_conv(h, W_res, K, Hp, seed, ffd_drop, stream_id, graph.row_offset)
                bias = bias + b_res
            plan = ops.NodeAttentionPlan(graphs=[graph], K=K, H=Hp, act=act_code, project_mode=mode, in_drop=ffd_drop,
                                         coef_drop=attn_drop, seed=seed, metapath_ids=[stream_id])
            z = ops.node_attention(plan, h, W, a1.unsqueeze(0), lay["b1"].unsqueeze(0), a2.unsqueeze(0),
                                   lay["b2"].unsqueeze(0), bias.unsqueeze(0), res)[:, 0, :]
            return z if Hp == H else z.view(-1, K, Hp)[:, :, :H].reshape(-1, K * H)

        h = x
        for l, (lay, (K, H)) in enumerate(zip(params.hidden, params.layer_dims)):   # :11-23
            h = layer(h, lay, K, H, act, l, project_mode if l == 0 else 0)
        K, C = params.out_heads, params.C
        out = layer(h, params.out, K, C, _lib.ACT_IDENTITY, len(params.hidden), 0)     # :25-29
        logits = out.view(-1, K, C).sum(1) / K                                        # :30
        return logits.unsqueeze(0)


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
        (multi-GPU context: ``dist.RowShard`` -- the lists hold this rank's rows of every meta-path -- or
        ``tiles.TileShard`` -- the lists hold only the meta-paths this rank owns, rows of its attention
        block, and the outputs are this rank's semantic rows), ``project_mode``, ``return_coef``.
        """
        from .tiles import TileShard
        tile = dist if isinstance(dist, TileShard) else None
        if tile is not None:
            if params is None:
                raise ValueError("tile sharding needs params= (the variables of ALL meta-paths, replicated)")
            if len(inputs_list) != len(tile.paths) or len(bias_mat_list) != len(tile.paths):
                raise ValueError(f"tile sharding: pass this rank's meta-paths {tile.paths} only")
            dist = tile.attn                                             # row exchange inside the meta-path's ranks
        gid = (lambda p: tile.paths[p]) if tile is not None else (lambda p: p)
        attn_drop, ffd_drop = float(attn_drop), float(ffd_drop)
        if not (0.0 <= attn_drop < 1.0 and 0.0 <= ffd_drop < 1.0):
            raise ValueError("attn_drop and ffd_drop are probabilities of dropping, in [0, 1)")
        if len(n_heads) < len(hid_units) + 1:
            raise ValueError("n_heads needs one entry per attention layer plus the output-layer entry")
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
                                             mp_att_size, device=dev, residual=residual)
                variables.set_default_store(params)
        act = ops.activation_code(activation)
        seed = None
        if attn_drop or ffd_drop:
            # training-mode dropout (ex_acm3025.py:185-186 feeds 0.6 / 0.6; 0.0 at evaluation): a fresh mask
            # set per call; the seed word lives on the device so a captured CUDA graph advances it too
            seed = ops.next_seed(params.drop_seed)                  # this call's own snapshot (fwd + bwd masks)

        coef_out = [None] * P
        groups = _group_by_input(xs)
        # tile sharding, single attention layer, one group: K-B stores Z straight into the semantic owners' buffers
        sink = tile.z_sink(K * H) if (tile is not None and len(hid_units) == 1 and len(groups) == 1) else None
        z_parts = []
        for gi, grp in enumerate(groups):
            plan = ops.NodeAttentionPlan(graphs=[graphs[p] for p in grp], K=K, H=H, act=act,
                                         project_mode=project_mode, dist=dist, want_coefs=return_coef,
                                         in_drop=ffd_drop, coef_drop=attn_drop, seed=seed,
                                         metapath_ids=[gid(p) for p in grp], slot=gi, z_sink=sink)
            if len(grp) == 1:
                p = gid(grp[0])
                W, a1, b1 = params.W[p], params.a1[p].unsqueeze(0), params.b1[p].unsqueeze(0)
                a2, b2, bias = params.a2[p].unsqueeze(0), params.b2[p].unsqueeze(0), params.bias[p].unsqueeze(0)
            else:
                ids = [gid(p) for p in grp]
                W = torch.cat([params.W[p] for p in ids], dim=1)
                a1 = torch.stack([params.a1[p] for p in ids]); b1 = torch.stack([params.b1[p] for p in ids])
                a2 = torch.stack([params.a2[p] for p in ids]); b2 = torch.stack([params.b2[p] for p in ids])
                bias = torch.stack([params.bias[p] for p in ids])
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

        # stacked layers (:48-57): every meta-path feeds its own concatenated heads to its next layer
        for l in range(1, len(hid_units)):
            Kl, Hl = int(n_heads[l]), int(hid_units[l])
            lay = params.deep[l - 1]
            use_res = residual and "W_res" in lay                           # layers.py:38-40 (no-op when widths agree, :42)
            nxt = []
            for p in range(P):
                h_old = multi_embed[:, p, :].contiguous()                   # :49
                q = gid(p)
                sid = l * params.P + q
                plan = ops.NodeAttentionPlan(graphs=[graphs[p]], K=Kl, H=Hl, act=act, project_mode=0, dist=dist,
                                             want_coefs=return_coef, in_drop=ffd_drop, coef_drop=attn_drop,
                                             seed=seed, metapath_ids=[sid], slot=1000 * l + p)
                res, bias_l = None, lay["bias"][q]
                if use_res:     # ret + conv1d(seq, H, 1) before the activation; every head reads its own dropped input
                    res = ops.residual_conv(h_old, lay["W_res"][q], Kl, Hl, seed, ffd_drop, sid, graphs[p].row_offset)
                    bias_l = bias_l + lay["b_res"][q]
                h = ops.node_attention(plan, h_old, lay["W"][q], lay["a1"][q].unsqueeze(0), lay["b1"][q].unsqueeze(0),
                                       lay["a2"][q].unsqueeze(0), lay["b2"][q].unsqueeze(0), bias_l.unsqueeze(0),
                                       res)[:, 0, :]
                if return_coef:
                    coef_out[p] = layers.EdgeCoefs(graphs[p], plan.coefs[0])
                nxt.append(h)
            multi_embed = torch.stack(nxt, dim=1)                           # (N,P,K_l*H_l)

        if tile is not None:
            # (rows of my attention block, my meta-paths, D) -> (my semantic rows, ALL meta-paths, D): one all-to-all
            multi_embed = tile.exchange_Z(multi_embed.contiguous(), pushed=bool(sink is not None and sink.used))

        final_embed, att_val = layers.SimpleAttLayer(                       # :61-63
            multi_embed, mp_att_size, time_major=False, return_alphas=True,
            params={"w_omega": params.w_omega, "b_omega": params.b_omega, "u_omega": params.u_omega},
            mode=semantic_mode, dist=tile if tile is not None else dist)

        out = []
        own = ops.dense_supported(final_embed.shape[1], params.Wc[0].shape[1])   # D <= 64, C <= 384: own kernels
        for i in range(n_heads[-1]):                                        # :66-68
            out.append(ops.dense(final_embed, params.Wc[i], params.bc[i]) if own
                       else torch.addmm(params.bc[i], final_embed, params.Wc[i]))
        logits = out[0] if len(out) == 1 else torch.stack(out).sum(0) / n_heads[-1]   # :72
        logits = logits.unsqueeze(0)                                        # :76
        if return_coef:
            return logits, final_embed, att_val, coef_out
        return logits, final_embed, att_val
