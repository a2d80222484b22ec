"""Generates tests/golden/ref_*.npz by EXECUTING THE REFERENCE'S OWN SOURCE, unmodified.

    python oracle/refrun/make_ref_golden.py          # needs /root/reference (this container only)

TEST INFRASTRUCTURE.  /root/reference/utils/process.py is pure numpy and runs as it is (one dead import
stubbed); /root/reference/utils/layers.py, models/gat.py and models/base_gattn.py run through the TF1 API
shim next to this file (oracle/refrun/tf1_shim.py), which implements each tf.* primitive they call with its
documented TF1 semantics on torch-CPU tensors.  The fixtures therefore hold what the reference's code
computes on seeded inputs and weights: outputs, loss, every gradient (torch autograd through the reference's
op chain = TF's graph autodiff), the variables after one `training()` step, attention coefficients.
tests/test_oracle.py asserts the oracle restatement (oracle/han_oracle.py) reproduces them to <= 1e-12 in
fp64; the -m gpu tests compare the CUDA path with these same files.  The fixtures travel to the GPU box;
/root/reference does not and is never read there.

Weights are handed to the reference in its own variable CREATION ORDER: per attn_head
call conv1d kernel (1,F,H), conv1d kernel+bias (1,H,1)/(1,), conv1d kernel+bias, BiasAdd biases (H,), and for a
residual head whose widths differ one more conv1d kernel+bias; then SimpleAttLayer's three tf.Variables; then
one tf.layers.dense per output head.  The shim checks kind and shape of every hand-over.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from han_b200 import synth  # noqa: E402
from oracle import han_oracle as O  # noqa: E402  (only its seeded parameter initialiser and layout helpers)
from oracle.refrun import tf1_shim as S  # noqa: E402
from tests.golden.trees import flatten_tree  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
REF = os.environ.get("HAN_REFERENCE_ROOT", "/root/reference")


# ---- creation-order hand-over of a head's variables ----------------------------------------------------------
def head_feed(hp, with_scores=True):
    """(kind, array) in the order utils/layers.py creates them (:20, :23, :24, :35, :40)."""
    W = hp["W"].detach().numpy()
    feed = [("conv1d/kernel", W[None])]
    slots = ["W"]
    if with_scores:
        H = W.shape[1]
        feed += [("conv1d/kernel", hp["a1"].detach().numpy().reshape(1, H, 1)),
                 ("conv1d/bias", hp["b1"].detach().numpy().reshape(1)),
                 ("conv1d/kernel", hp["a2"].detach().numpy().reshape(1, H, 1)),
                 ("conv1d/bias", hp["b2"].detach().numpy().reshape(1))]
        slots += ["a1", "b1", "a2", "b2"]
    feed.append(("BiasAdd/biases", hp["bias"].detach().numpy()))
    slots.append("bias")
    if "W_res" in hp:
        feed += [("conv1d/kernel", hp["W_res"].detach().numpy()[None]), ("conv1d/bias", hp["b_res"].detach().numpy())]
        slots += ["W_res", "b_res"]
    return feed, slots


def han_feed(params, P, hid_units, n_heads):
    """Creation order of HeteGAT_multi / HeteGAT .inference (models/gat.py:39-68): meta-path major, layer, head."""
    feed, where = [], []
    for p in range(P):
        for layer in range(len(hid_units)):
            for k in range(n_heads[layer]):
                hp = O.head_params(params, p, k, layer=layer)
                f, s = head_feed(hp)
                feed += f
                where += [(layer, p, k, name) for name in s]
    feed += [("Variable", params["w_omega"].detach().numpy()), ("Variable", params["b_omega"].detach().numpy()),
             ("Variable", params["u_omega"].detach().numpy())]
    where += [(None, None, None, n) for n in ("w_omega", "b_omega", "u_omega")]
    for i in range(n_heads[-1]):
        feed += [("dense/kernel", params["Wc"][i].detach().numpy()), ("dense/bias", params["bc"][i].detach().numpy())]
        where += [(None, None, i, "Wc"), (None, None, i, "bc")]
    return feed, where


def han_unfeed(values, where, params, hid_units, n_heads):
    """Per-variable tensors (gradients, updated variables) in creation order -> the concatenated parameter layout."""
    out = O.params_to(params, torch.float64)
    for t in O.flat_params(out):
        t.zero_()
    for v, (layer, p, k, name) in zip(values, where):
        v = v.detach()
        if layer is None:
            if name in ("Wc", "bc"):
                out[name][k].copy_(v)
            else:
                out[name].copy_(v)
            continue
        dst = out if layer == 0 else out["deep"][layer - 1]
        H = hid_units[layer]
        sl = slice(k * H, (k + 1) * H)
        if name in ("W", "W_res"):
            dst[name][p][:, sl] = v[0]
        elif name in ("bias", "b_res"):
            dst[name][p][sl] = v
        elif name in ("a1", "a2"):
            dst[name][p][k] = v.reshape(-1)
        else:
            dst[name][p][k] = v.reshape(())
    return out


def _common(cfg):
    out = {"X": cfg.X, "labels": cfg.labels, "train_mask": cfg.train_mask}
    for i, (indptr, indices) in enumerate(cfg.csr()):
        out[f"indptr{i}"], out[f"indices{i}"] = indptr, indices
    return out


def _old_format_params(params, prefix):
    from tests.golden import make_golden
    return make_golden.flatten_params(params, prefix)


# ---- cases ----------------------------------------------------------------------------------------------------
def case_adj_to_bias(process):
    """utils/process.py:14-25 run as shipped: known-answer inputs of SURVEY 8(c) + seeded random ones."""
    out = {}
    n = 6
    adj = np.zeros((1, n, n))
    adj[0, 0, 1] = 3.0; adj[0, 2, 3] = -1.0; adj[0, 4, 4] = -1.0
    cases = {"known": (adj, [n], 1)}
    rng = np.random.default_rng(7)
    a = (rng.random((1, 40, 40)) < 0.08).astype(np.float64) * rng.integers(1, 4, size=(1, 40, 40))
    a[0, 3, 7] = -2.0
    cases["weighted_nh1"] = (a, [40], 1)
    cases["weighted_nh2"] = (a, [40], 2)
    cases["nh3"] = ((rng.random((1, 33, 33)) < 0.05).astype(np.float64), [33], 3)
    cfg = synth.tiny(seed=21, n=96, f=40, p=2, c=3, deg=6.0)
    cases["pap_minus_I"] = (cfg.adjs()[0], [cfg.N], 1)
    two = np.concatenate([cfg.adjs()[0], cfg.adjs()[1]], 0)
    cases["two_graphs"] = (two, [cfg.N, cfg.N], 1)
    cases["partial_sizes"] = (a, [25], 1)         # entries outside sizes[g]^2 keep their raw value (process.py:21-24)
    for name, (adj, sizes, nhood) in cases.items():
        out[f"{name}/adj"] = adj
        out[f"{name}/sizes"] = np.asarray(sizes)
        out[f"{name}/nhood"] = np.asarray(nhood)
        out[f"{name}/bias"] = process.adj_to_bias(adj, sizes, nhood)
    return out


def run_han_multi(gat, base, cfg, heads, hid, att, seed, deep=(), residual=False, dtype=torch.float64):
    """HeteGAT_multi.inference -> masked_softmax_cross_entropy -> training() (L2 + one Adam step), as the driver
    wires them (ex_acm3025.py:139-152), through the reference's own functions."""
    hid_units = [hid] + [h for (_, h) in deep]
    n_heads = [heads] + [k for (k, _) in deep] + [1]
    params = O.init_params(np.random.default_rng(seed), [cfg.F] * cfg.P, cfg.C, hid=hid, heads=heads, mp_att_size=att,
                           deep=deep, residual=residual)
    S.STORE.reset(dtype=dtype)
    feed, where = han_feed(params, cfg.P, hid_units, n_heads)
    S.STORE.feed(feed)
    X = S.Tensor(torch.from_numpy(cfg.X).to(dtype)[None])
    _, process, _, _ = REFMODS
    biases = [S.Tensor(torch.from_numpy(process.adj_to_bias(a, [cfg.N], nhood=1)).to(dtype)) for a in cfg.adjs()]
    logits, final_embed, att_val = gat.HeteGAT_multi.inference(
        [X] * cfg.P, cfg.C, cfg.N, True, 0.0, 0.0, bias_mat_list=biases, hid_units=hid_units, n_heads=n_heads,
        residual=residual, activation=TF.nn.elu, mp_att_size=att)
    assert not S.STORE.queue, "not every fed variable was created"
    names = [v.name for v in S.STORE.variables]
    log_resh = S.reshape(logits, [-1, cfg.C])                                       # ex_acm3025.py:146
    lab_resh = S.Tensor(torch.from_numpy(cfg.labels).to(dtype))
    msk = S.Tensor(torch.from_numpy(cfg.train_mask.astype(np.int32)))
    loss = base.BaseGAttN.masked_softmax_cross_entropy(log_resh, lab_resh, msk)   # :149
    acc = base.BaseGAttN.masked_accuracy(log_resh, lab_resh, msk)                 # :150
    # total = loss + lossL2 as training() builds it (base_gattn.py:14-22); read it back from the gradients' source
    vars_ = S.trainable_variables()
    lossL2 = sum((v.t * v.t).sum() / 2 for v in vars_) * 0.001
    base.BaseGAttN.training(loss, 0.005, 0.001)                                    # :152  -> grads + one Adam step
    grads = han_unfeed(S.STORE.last_grads, where, params, hid_units, n_heads)
    after = han_unfeed([v.t for v in S.STORE.variables], where, params, hid_units, n_heads)
    out = _common(cfg)
    out.update({"meta": np.array([cfg.N, cfg.F, cfg.P, cfg.C, heads, hid, att, 0], dtype=np.int64),
                "logits": logits.numpy(), "final_embed": final_embed.numpy(), "att_val": att_val.numpy(),
                "ce": loss.numpy(), "acc": acc.numpy(), "total": (loss.t + lossL2).detach().numpy(),
                "tf_names": np.array(names)})
    if deep:
        out["deep"] = np.asarray(deep, dtype=np.int64)
        out["residual"] = np.asarray(int(residual))
        out.update(flatten_tree(params, "p"))
        out.update(flatten_tree(grads, "g"))
        out.update(flatten_tree(after, "a"))
    else:
        out.update(_old_format_params(params, "p_"))
        out.update(_old_format_params(grads, "g_"))
        out.update(_old_format_params(after, "a_"))
    return out


def case_han_multi_p2_k8h8(mods):
    return run_han_multi(mods[2], mods[3], synth.tiny(seed=21, n=96, f=40, p=2, c=3, deg=6.0), 8, 8, 128, 121)


def case_han_multi_p3_k4h8(mods):
    return run_han_multi(mods[2], mods[3], synth.tiny(seed=22, n=70, f=33, p=3, c=4, deg=4.0), 4, 8, 64, 122)


def case_han_multi_degenerate(mods):
    cfg = synth.tiny(seed=23, n=50, f=12, p=2, c=3, deg=5.0, binary=True)
    m0, m1 = cfg.masks
    m0[3, :] = False; m0[3, 3] = True
    m0[7, :] = True
    m1[:, 11] = True
    m1[20, :] = False; m1[20, 5] = True
    return run_han_multi(mods[2], mods[3], cfg, 8, 8, 128, 123)


def case_han_multi_fp32(mods):
    """The same graph as p2_k8h8 in fp32: what the reference computes at its own precision."""
    return run_han_multi(mods[2], mods[3], synth.tiny(seed=21, n=96, f=40, p=2, c=3, deg=6.0), 8, 8, 128, 121,
                         dtype=torch.float32)


def case_han_multi_stacked_residual(mods):
    """hid_units=[8,8,4], n_heads=[4,2,4,1], residual=True (models/gat.py:48-57, utils/layers.py:38-42): layer 1 maps
    32 -> 8 (widths differ: conv1d residual), layer 2 maps 16 -> 4 (differ)."""
    return run_han_multi(mods[2], mods[3], synth.tiny(seed=24, n=64, f=20, p=2, c=3, deg=5.0), 4, 8, 32, 124,
                         deep=((2, 8), (4, 4)), residual=True)


def case_hetegat_coefs(mods):
    """HeteGAT.inference(inputs, ..., return_coef=True) (models/gat.py:132-203): shared inputs, per meta-path the
    head-averaged dense attention matrix."""
    gat = mods[2]
    cfg = synth.tiny(seed=25, n=60, f=18, p=2, c=3, deg=5.0)
    heads, hid, att = 4, 8, 32
    params = O.init_params(np.random.default_rng(125), [cfg.F] * cfg.P, cfg.C, hid=hid, heads=heads, mp_att_size=att)
    S.STORE.reset()
    feed, _ = han_feed(params, cfg.P, [hid], [heads, 1])
    S.STORE.feed(feed)
    X = S.Tensor(torch.from_numpy(cfg.X).double()[None])
    biases = [S.Tensor(torch.from_numpy(REFMODS[1].adj_to_bias(a, [cfg.N], nhood=1))) for a in cfg.adjs()]
    logits, fe, av, coef_list = gat.HeteGAT.inference(X, cfg.C, cfg.N, False, 0.0, 0.0, biases, [hid], [heads, 1],
                                                      activation=TF.nn.elu, mp_att_size=att,
                                                      return_coef=True)
    out = _common(cfg)
    out.update({"meta": np.array([cfg.N, cfg.F, cfg.P, cfg.C, heads, hid, att, 0], dtype=np.int64),
                "logits": logits.numpy(), "final_embed": fe.numpy(), "att_val": av.numpy()})
    for i, c in enumerate(coef_list):
        out[f"coef{i}"] = c.numpy()
    out.update(_old_format_params(params, "p_"))
    return out


def _gat_feed(params, hid_units, n_heads, nb_classes, residual):
    feed, where = [], []
    F_in = params["hidden"][0]["W"].shape[0]
    for l, H in enumerate(hid_units):
        lay = params["hidden"][l]
        for k in range(n_heads[l]):
            sl = slice(k * H, (k + 1) * H)
            hp = {"W": lay["W"][:, sl], "a1": lay["a1"][k], "b1": lay["b1"][k], "a2": lay["a2"][k], "b2": lay["b2"][k],
                  "bias": lay["bias"][sl]}
            if "W_res" in lay:
                hp["W_res"], hp["b_res"] = lay["W_res"][:, sl], lay["b_res"][sl]
            f, s = head_feed(hp)
            feed += f
            where += [("hidden", l, k, H, n) for n in s]
    lay = params["out"]
    H = nb_classes
    for k in range(n_heads[-1]):
        sl = slice(k * H, (k + 1) * H)
        hp = {"W": lay["W"][:, sl], "a1": lay["a1"][k], "b1": lay["b1"][k], "a2": lay["a2"][k], "b2": lay["b2"][k],
              "bias": lay["bias"][sl]}
        f, s = head_feed(hp)
        feed += f
        where += [("out", None, k, H, n) for n in s]
    return feed, where


def _gat_unfeed(values, where, params):
    out = {"hidden": [{k: torch.zeros_like(v) for k, v in lay.items()} for lay in params["hidden"]],
           "out": {k: torch.zeros_like(v) for k, v in params["out"].items()}}
    for v, (grp, l, k, H, name) in zip(values, where):
        v = v.detach()
        dst = out["hidden"][l] if grp == "hidden" else out["out"]
        sl = slice(k * H, (k + 1) * H)
        if name in ("W", "W_res"):
            dst[name][:, sl] = v[0]
        elif name in ("bias", "b_res"):
            dst[name][sl] = v
        elif name in ("a1", "a2"):
            dst[name][k] = v.reshape(-1)
        else:
            dst[name][k] = v.reshape(())
    return out


def case_gat(mods, hid_units=(8, 8), n_heads=(2, 2, 2), residual=True, classes=7, seed=131):
    """GAT.inference (models/gat.py:8-32) + masked CE, gradients of every variable."""
    gat, base = mods[2], mods[3]
    cfg = synth.tiny(seed=seed, n=90, f=24, p=1, c=classes, deg=6.0)
    params = O.init_gat_params(np.random.default_rng(seed + 1), cfg.F, cfg.C, hid_units, n_heads, residual=residual)
    S.STORE.reset()
    feed, where = _gat_feed(params, hid_units, n_heads, cfg.C, residual)
    S.STORE.feed(feed)
    X = S.Tensor(torch.from_numpy(cfg.X).double()[None])
    bias = S.Tensor(torch.from_numpy(REFMODS[1].adj_to_bias(cfg.adjs()[0], [cfg.N], nhood=1)))
    logits = gat.GAT.inference(X, cfg.C, cfg.N, False, 0.0, 0.0, bias, list(hid_units), list(n_heads),
                               activation=TF.nn.elu, residual=residual)
    assert not S.STORE.queue
    lab = S.Tensor(torch.from_numpy(cfg.labels).double())
    msk = S.Tensor(torch.from_numpy(cfg.train_mask.astype(np.int32)))
    loss = base.BaseGAttN.masked_softmax_cross_entropy(S.reshape(logits, [-1, cfg.C]), lab, msk)
    gr = torch.autograd.grad(loss.t, [v.t for v in S.STORE.variables])
    grads = _gat_unfeed(gr, where, params)
    out = _common(cfg)
    out.update({"meta": np.array([cfg.N, cfg.F, 1, cfg.C], dtype=np.int64), "hid_units": np.asarray(hid_units),
                "n_heads": np.asarray(n_heads), "residual": np.asarray(int(residual)),
                "logits": logits.numpy(), "ce": loss.numpy()})
    out.update(flatten_tree(params, "p"))
    out.update(flatten_tree(grads, "g"))
    return out


def _single_head_inputs(seed, n=70, f=22, h=8):
    cfg = synth.tiny(seed=seed, n=n, f=f, p=1, deg=6.0)
    rng = np.random.default_rng(seed + 1)
    lim = np.sqrt(6.0 / (f + h))
    hp = {"W": torch.from_numpy(rng.uniform(-lim, lim, (f, h))), "a1": torch.from_numpy(rng.normal(size=h)),
          "b1": torch.tensor(0.05, dtype=torch.float64), "a2": torch.from_numpy(rng.normal(size=h)),
          "b2": torch.tensor(-0.03, dtype=torch.float64), "bias": torch.from_numpy(rng.normal(0, 0.1, h)),
          "W_res": torch.from_numpy(rng.uniform(-lim, lim, (f, h))), "b_res": torch.from_numpy(rng.normal(0, 0.1, h))}
    return cfg, hp


def _head_case(fn_name, layers, seed, with_scores, residual, **call):
    cfg, hp = _single_head_inputs(seed)
    if not residual:
        hp.pop("W_res"); hp.pop("b_res")
    S.STORE.reset()
    feed, slots = head_feed(hp, with_scores=with_scores)
    S.STORE.feed(feed)
    X = S.Tensor(torch.from_numpy(cfg.X).double()[None])
    bias = S.Tensor(torch.from_numpy(REFMODS[1].adj_to_bias(cfg.adjs()[0], [cfg.N], nhood=1)))
    r = getattr(layers, fn_name)(X, 8, bias, TF.nn.elu, residual=residual, **call)
    assert not S.STORE.queue
    coefs = None
    if isinstance(r, tuple):
        r, coefs = r
    g = torch.from_numpy(np.random.default_rng(seed + 2).normal(size=tuple(r.t.shape)))
    gr = torch.autograd.grad((r.t * g).sum(), [v.t for v in S.STORE.variables])
    out = _common(cfg)
    out.update({"out": r.numpy(), "cot": g.numpy()})
    if coefs is not None:
        out["coefs"] = coefs.numpy()
    for s, v in zip(slots, S.STORE.variables):
        out[f"p/{s}"] = hp[s].numpy()
    for s, v in zip(slots, gr):
        out[f"g/{s}"] = v.numpy().reshape(hp[s].shape)
    return out


def case_attn_head(mods):
    """One attn_head call (utils/layers.py:7-46) with return_coef=True and the residual conv."""
    return _head_case("attn_head", mods[0], 111, True, True, return_coef=True)


def case_attn_head_const_1(mods):
    """attn_head_const_1 (utils/layers.py:49-81) with the residual conv."""
    return _head_case("attn_head_const_1", mods[0], 101, False, True)


def _dropout_case(layers, seed, residual):
    cfg, hp = _single_head_inputs(seed, n=60, f=16)
    if not residual:
        hp.pop("W_res"); hp.pop("b_res")
    keep = 0.4
    N, F, H = cfg.N, cfg.F, 8
    rng = np.random.default_rng(seed + 1)
    mx = (rng.random((1, N, F)) < keep)
    mc = (rng.random((1, N, N)) < keep)
    ms = (rng.random((1, N, H)) < keep)
    S.STORE.reset()
    feed, slots = head_feed(hp)
    S.STORE.feed(feed)
    S.STORE.feed_masks([mx, mc, ms])
    X = S.Tensor(torch.from_numpy(cfg.X).double()[None])
    bias = S.Tensor(torch.from_numpy(REFMODS[1].adj_to_bias(cfg.adjs()[0], [cfg.N], nhood=1)))
    r, coefs = layers.attn_head(X, 8, bias, TF.nn.elu, in_drop=0.6, coef_drop=0.6, residual=residual, return_coef=True)
    assert not S.STORE.masks and [c[0] for c in S.STORE.dropout_calls] == [(1, N, F), (1, N, N), (1, N, H)]
    assert not S.STORE.queue
    g = torch.from_numpy(np.random.default_rng(seed + 2).normal(size=tuple(r.t.shape)))
    gr = torch.autograd.grad((r.t * g).sum(), [v.t for v in S.STORE.variables])
    out = _common(cfg)
    out.update({"out": r.numpy(), "cot": g.numpy(), "coefs": coefs.numpy(), "mask_x": mx, "mask_coef": mc, "mask_s": ms,
                "keep": np.asarray(keep)})
    for s in slots:
        out[f"p/{s}"] = hp[s].numpy()
    for s, v in zip(slots, gr):
        out[f"g/{s}"] = v.numpy().reshape(hp[s].shape)
    return out


def case_attn_head_dropout(mods):
    """attn_head with in_drop = coef_drop = 0.6 (utils/layers.py:18-19,29-32), the three keep masks handed to
    tf.nn.dropout in call order (input X, coefficients, projected features) so another implementation given the SAME
    masks can be compared exactly."""
    return _dropout_case(mods[0], 141, False)


def case_attn_head_dropout_residual(mods):
    """The training configuration with residual=True: the residual conv1d (utils/layers.py:40) reads the DROPPED
    input of :19, not the clean one."""
    return _dropout_case(mods[0], 145, True)


def case_sp_attn_head(mods):
    """sp_attn_head (utils/layers.py:85-127) on a WEIGHTED sparse adjacency: the logits are
    adj_ij * f1_i + adj_ij * f2_j over the stored entries (:95-96), leaky_relu, tf.sparse_softmax over each row's
    stored entries, sparse @ dense.  Weighted entries are what distinguishes it from attn_head."""
    layers = mods[0]
    cfg, hp = _single_head_inputs(151, n=66, f=20)
    hp.pop("W_res"); hp.pop("b_res")
    m = cfg.masks[0]
    rng = np.random.default_rng(152)
    w = np.where(m, rng.uniform(0.25, 2.0, size=m.shape), 0.0)
    rows, cols = np.nonzero(m)
    idx = np.stack([np.zeros_like(rows), rows, cols], 1)
    out = _common(cfg)
    for tag, vals in (("binary", np.ones(len(rows))), ("weighted", w[rows, cols])):
        S.STORE.reset()
        feed, slots = head_feed(hp)
        S.STORE.feed(feed)
        X = S.Tensor(torch.from_numpy(cfg.X).double()[None])
        adj = S.SparseTensor(idx, torch.from_numpy(vals), [1, cfg.N, cfg.N])
        r = layers.sp_attn_head(X, 8, adj, TF.nn.elu, cfg.N)
        g = torch.from_numpy(np.random.default_rng(153).normal(size=tuple(r.t.shape)))
        gr = torch.autograd.grad((r.t * g).sum(), [v.t for v in S.STORE.variables])
        out[f"{tag}/values"] = vals
        out[f"{tag}/out"] = r.numpy()
        out[f"{tag}/cot"] = g.numpy()
        for s, v in zip(slots, gr):
            out[f"{tag}/g/{s}"] = v.numpy().reshape(hp[s].shape)
    for s in slots:
        out[f"p/{s}"] = hp[s].numpy()
    return out


def case_semantic(mods):
    """SimpleAttLayer alone (utils/layers.py:132-164), return_alphas=True, with its gradients."""
    layers = mods[0]
    rng = np.random.default_rng(161)
    N, P, D, A = 77, 3, 64, 128
    Z = torch.from_numpy(rng.normal(size=(N, P, D))).requires_grad_(True)
    sp = {"w_omega": rng.normal(0, 0.1, (D, A)), "b_omega": rng.normal(0, 0.1, A), "u_omega": rng.normal(0, 0.1, A)}
    S.STORE.reset()
    S.STORE.feed([("Variable", sp["w_omega"]), ("Variable", sp["b_omega"]), ("Variable", sp["u_omega"])])
    o, al = layers.SimpleAttLayer(S.Tensor(Z), A, time_major=False, return_alphas=True)
    g = torch.from_numpy(rng.normal(size=(N, D)))
    gr = torch.autograd.grad((o.t * g).sum(), [Z] + [v.t for v in S.STORE.variables])
    out = {"Z": Z.detach().numpy(), "out": o.numpy(), "alphas": al.numpy(), "cot": g.numpy(), "g/Z": gr[0].numpy()}
    for k, v in zip(("w_omega", "b_omega", "u_omega"), gr[1:]):
        out[f"p/{k}"] = sp[k]
        out[f"g/{k}"] = v.numpy()
    out["tf_names"] = np.array([v.name for v in S.STORE.variables])
    return out


CASES = {
    "ref_han_multi_p2_k8h8": case_han_multi_p2_k8h8,
    "ref_han_multi_p3_k4h8": case_han_multi_p3_k4h8,
    "ref_han_multi_degenerate": case_han_multi_degenerate,
    "ref_han_multi_fp32": case_han_multi_fp32,
    "ref_han_multi_stacked_residual": case_han_multi_stacked_residual,
    "ref_hetegat_coefs": case_hetegat_coefs,
    "ref_gat": case_gat,
    "ref_attn_head": case_attn_head,
    "ref_attn_head_const_1": case_attn_head_const_1,
    "ref_attn_head_dropout": case_attn_head_dropout,
    "ref_attn_head_dropout_residual": case_attn_head_dropout_residual,
    "ref_sp_attn_head": case_sp_attn_head,
    "ref_semantic": case_semantic,
}

REFMODS = None
TF = None


def main():
    global REFMODS, TF
    TF = S.install()
    assert os.path.isdir(REF), f"{REF} not present: the fixtures are generated in the build container only"
    REFMODS = S.import_reference(REF)
    arrays = case_adj_to_bias(REFMODS[1])
    np.savez_compressed(os.path.join(GOLDEN, "ref_adj_to_bias.npz"), **arrays)
    print("ref_adj_to_bias", sum(a.nbytes for a in arrays.values()) // 1024, "KiB")
    import contextlib
    import io
    for name, fn in CASES.items():
        with contextlib.redirect_stdout(io.StringIO()):      # the reference prints 'de' (models/gat.py:74)
            arrays = fn(REFMODS)
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **arrays)
        print(name, sum(np.asarray(a).nbytes for a in arrays.values()) // 1024, "KiB")


if __name__ == "__main__":
    main()
