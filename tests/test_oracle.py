"""Pins for the CPU oracle (the reference ships no tests: "parity unpinned" at the TF boundary).
Known-answer cases of SURVEY.md section 8(c), internal consistency (dense vs edge-list twin, loop vs
vectorised adj_to_bias), fp64 gradcheck, and the committed golden fixtures."""
import os

import numpy as np
import pytest
import torch

from han_b200 import synth
from oracle import han_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _cfg_and_params(seed=3, n=48, f=20, p=2, heads=4, hid=4, att=16):
    cfg = synth.tiny(seed=seed, n=n, f=f, p=p, c=3, deg=5.0)
    rng = np.random.default_rng(seed + 100)
    params = O.init_params(rng, [cfg.F] * cfg.P, cfg.C, hid=hid, heads=heads, mp_att_size=att)
    return cfg, params


# ---- adj_to_bias -------------------------------------------------------------------------------
def test_adj_to_bias_loop_equals_vectorised():
    rng = np.random.default_rng(0)
    adj = (rng.random((1, 30, 30)) < 0.1).astype(np.float64) * rng.integers(1, 4, size=(1, 30, 30))
    adj[0, 3, 7] = -2.0  # negative entry stays un-thresholded (process.py:23 only rewrites > 0)
    for nhood in (1, 2):
        a = O.adj_to_bias(adj, [30], nhood)
        b = O.adj_to_bias_loop(adj, [30], nhood)
        assert np.array_equal(a, b)


def test_adj_to_bias_known_answers():
    n = 6
    adj = np.zeros((1, n, n))
    adj[0, 0, 1] = 3.0          # weighted count -> same mask as 1
    adj[0, 2, 3] = -1.0         # negative -> bias below -1e9, still masked
    adj[0, 4, 4] = -1.0         # 'PAP - I' with PAP_ii = 0 -> no self-loop
    b = O.adj_to_bias(adj, [n], 1)[0]
    assert b[0, 1] == 0.0 and b[0, 0] == 0.0
    assert b[2, 3] == -2e9
    assert b[4, 4] == -1e9 and np.all(b[4] <= -1e9)        # row 4 has no edge at all
    assert b[1, 2] == -1e9
    indptr, indices = O.bias_to_csr(b)
    assert indptr.tolist() == [0, 2, 3, 4, 5, 5, 6]
    assert indices.tolist() == [0, 1, 1, 2, 3, 5]


def test_adj_to_bias_nhood2_is_reachability():
    rng = np.random.default_rng(1)
    adj = (rng.random((1, 25, 25)) < 0.08).astype(np.float64)
    b = O.adj_to_bias(adj, [25], 2)[0]
    reach = np.linalg.matrix_power(adj[0] + np.eye(25), 2) > 0
    assert np.array_equal(b == 0, reach)


# ---- attn_head known answers ----------------------------------------------------------------------
def _head_params(rng, F, H, dtype=torch.float64):
    return {"W": torch.from_numpy(rng.normal(size=(F, H))).to(dtype),
            "a1": torch.from_numpy(rng.normal(size=(H,))).to(dtype), "b1": torch.tensor(0.3, dtype=dtype),
            "a2": torch.from_numpy(rng.normal(size=(H,))).to(dtype), "b2": torch.tensor(-0.2, dtype=dtype),
            "bias": torch.from_numpy(rng.normal(size=(H,))).to(dtype)}


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_all_masked_row_is_uniform(dtype):
    # a row whose bias is all -1e9 gets uniform attention 1/N over ALL nodes
    rng = np.random.default_rng(2)
    n, F, H = 12, 5, 4
    hp = _head_params(rng, F, H, dtype)
    x = torch.from_numpy(rng.normal(size=(1, n, F))).to(dtype)
    bias = torch.full((1, n, n), -1e9, dtype=dtype)
    bias[0, torch.arange(n), torch.arange(n)] = 0.0
    bias[0, 5, :] = -1e9
    out, coefs = O.attn_head(x, H, bias, O.identity, hp, return_coef=True)
    S = x[0] @ hp["W"]
    if dtype == torch.float32:
        # the reference's precision: ulp(1e9) = 64 absorbs every logit -> exactly uniform
        assert torch.allclose(coefs[0, 5], torch.full((n,), 1.0 / n, dtype=dtype), rtol=1e-6)
        assert torch.allclose(out[0, 5], S.mean(0) + hp["bias"], rtol=1e-5, atol=1e-6)
    else:
        # fp64 keeps the logits to ~1e-7 under the -1e9 offset: softmax over ALL nodes, not uniform.
        # (this fp32 artefact is why such rows are checked against the fp32 oracle only)
        f1 = S @ hp["a1"] + hp["b1"]
        f2 = S @ hp["a2"] + hp["b2"]
        ref = torch.softmax(torch.nn.functional.leaky_relu(f1[5] + f2, 0.2), 0)
        assert torch.allclose(coefs[0, 5], ref, rtol=1e-5)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_single_neighbour_alpha_is_one_and_masked_is_exact_zero(dtype):
    rng = np.random.default_rng(3)
    n, F, H = 9, 4, 4
    hp = _head_params(rng, F, H, dtype)
    x = torch.from_numpy(rng.normal(size=(1, n, F))).to(dtype)
    bias = torch.full((1, n, n), -1e9, dtype=dtype)
    bias[0, torch.arange(n), torch.arange(n)] = 0.0       # self-loop only
    out, coefs = O.attn_head(x, H, bias, O.elu, hp, return_coef=True)
    assert torch.equal(coefs[0], torch.eye(n, dtype=dtype))  # alpha = 1 exactly, masked exactly 0
    S = x[0] @ hp["W"]
    assert torch.allclose(out[0], torch.nn.functional.elu(S + hp["bias"]))


def test_dense_equals_edge_list_twin_fp64():
    cfg, params = _cfg_and_params()
    X = torch.from_numpy(cfg.X).double()
    biases = [torch.from_numpy(O.adj_to_bias(a, [cfg.N], 1)) for a in cfg.adjs()]
    lo, fe, av = O.HeteGAT_multi_inference([X.unsqueeze(0)] * cfg.P, cfg.C, cfg.N, False, 0.0, 0.0, biases,
                                           [4], [4, 1], params, mp_att_size=16)
    csr = [O.bias_to_csr(b.numpy()) for b in biases]
    assert all(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) for a, b in zip(csr, cfg.csr()))
    lo2, fe2, av2 = O.inference_edges([X] * cfg.P, csr, params, [4, 1], [4], mp_att_size=16)
    assert (lo - lo2).abs().max() < 1e-12 and (fe - fe2).abs().max() < 1e-12 and (av - av2).abs().max() < 1e-12


def test_semantic_p1_beta_is_one_and_paper_mode():
    rng = np.random.default_rng(4)
    Z = torch.from_numpy(rng.normal(size=(10, 1, 8)))
    sp = {"w_omega": torch.from_numpy(rng.normal(size=(8, 6))), "b_omega": torch.from_numpy(rng.normal(size=6)),
          "u_omega": torch.from_numpy(rng.normal(size=6))}
    out, al = O.SimpleAttLayer(Z, 6, sp, return_alphas=True)
    assert torch.equal(al, torch.ones(10, 1, dtype=torch.float64)) and torch.equal(out, Z[:, 0])
    Z3 = torch.from_numpy(rng.normal(size=(10, 3, 8)))
    out_p, al_p = O.SimpleAttLayer(Z3, 6, sp, return_alphas=True, mode="paper")
    v = torch.tanh(Z3 @ sp["w_omega"] + sp["b_omega"]) @ sp["u_omega"]
    beta = torch.softmax(v.mean(0), 0)
    assert torch.allclose(al_p, beta.expand(10, 3)) and torch.allclose(out_p, (Z3 * beta[None, :, None]).sum(1))


def test_masked_cross_entropy_matches_definition():
    rng = np.random.default_rng(5)
    logits = torch.from_numpy(rng.normal(size=(20, 3)))
    y = torch.nn.functional.one_hot(torch.from_numpy(rng.integers(0, 3, 20)), 3).double()
    m = torch.from_numpy((rng.random(20) < 0.4).astype(np.float64))
    got = O.masked_softmax_cross_entropy(logits, y, m)
    ref = torch.nn.functional.cross_entropy(logits[m.bool()], y[m.bool()].argmax(1))  # sum_masked / #masked
    assert torch.allclose(got, ref)


def test_adam_tf1_first_step():
    p, g = torch.tensor([1.0, -2.0]), torch.tensor([0.5, -0.25])
    p1, m, v = O.adam_step_tf1(p, g, torch.zeros(2), torch.zeros(2), 1, lr=0.005)
    # t=1: m_hat/sqrt(v_hat) = sign(g) up to eps -> step of lr
    assert torch.allclose(p1, p - 0.005 * torch.sign(g), atol=1e-6)


# ---- gradients ---------------------------------------------------------------------------------
def test_gradcheck_attn_head_and_semantic():
    rng = np.random.default_rng(6)
    n, F, H = 7, 3, 4
    hp = {k: v.requires_grad_(True) for k, v in _head_params(rng, F, H).items()}
    x = torch.from_numpy(rng.normal(size=(1, n, F)))
    mask = rng.random((n, n)) < 0.4
    np.fill_diagonal(mask, True)
    bias = torch.from_numpy(np.where(mask, 0.0, -1e9))[None]
    keys = list(hp)

    def f(*vals):
        return O.attn_head(x, H, bias, O.elu, dict(zip(keys, vals)))
    assert torch.autograd.gradcheck(f, tuple(hp[k] for k in keys), eps=1e-6, atol=1e-5)

    Z = torch.from_numpy(rng.normal(size=(5, 3, 4))).requires_grad_(True)
    sp = [torch.from_numpy(rng.normal(size=s)).requires_grad_(True) for s in ((4, 6), (6,), (6,))]
    for mode in ("reference", "paper"):
        def g(Z, w, b, u):
            return O.SimpleAttLayer(Z, 6, {"w_omega": w, "b_omega": b, "u_omega": u}, mode=mode)
        assert torch.autograd.gradcheck(g, (Z, *sp), eps=1e-6, atol=1e-5)


def test_fp32_oracle_close_to_fp64_gold():
    cfg, params = _cfg_and_params(seed=11, n=64, f=24)
    outs = {}
    for dt in (torch.float64, torch.float32):
        p = O.params_to(params, dt)
        X = torch.from_numpy(cfg.X).to(dt).unsqueeze(0)
        biases = [torch.from_numpy(O.adj_to_bias(a, [cfg.N], 1)).to(dt) for a in cfg.adjs()]
        outs[dt] = O.HeteGAT_multi_inference([X] * cfg.P, cfg.C, cfg.N, False, 0.0, 0.0, biases, [4], [4, 1], p,
                                             mp_att_size=16)
    for a, b in zip(outs[torch.float32], outs[torch.float64]):
        assert (a.double() - b).abs().max() / b.abs().max() < 5e-6


# ---- golden fixtures -----------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["tiny_p2_k8h8", "tiny_p3_k4h8_paper", "degenerate_rows"])
def test_oracle_reproduces_golden(name):
    from tests.golden import make_golden
    path = os.path.join(GOLDEN, name + ".npz")
    assert os.path.exists(path), "run python tests/golden/make_golden.py"
    stored = np.load(path)
    fresh = make_golden.CASES[name]()
    for k in stored.files:
        a, b = stored[k], fresh[k]
        if a.dtype.kind in "iu":
            assert np.array_equal(a, b), k
        else:
            assert np.allclose(a, b, rtol=1e-12, atol=1e-14), k


# ---- the widened rows: restatements pinned by closed forms / compositions -----------
def test_const_1_head_is_the_neighbour_mean():
    """utils/layers.py:49-81: logits = the 0/1 adjacency -> after the -1e9 mask every neighbour weighs 1/deg."""
    cfg = synth.tiny(seed=21, n=40, f=12, p=1, deg=5.0)
    rng = np.random.default_rng(22)
    hp = {"W": torch.from_numpy(rng.normal(size=(cfg.F, 4))), "bias": torch.from_numpy(rng.normal(size=4))}
    X = torch.from_numpy(cfg.X).double()[None]
    bias = torch.from_numpy(O.adj_to_bias(cfg.adjs()[0], [cfg.N], 1))
    out = O.attn_head_const_1(X, 4, bias, O.identity, hp)
    m = torch.from_numpy(cfg.masks[0]).double()
    ref = (m / m.sum(1, keepdim=True)) @ (X[0] @ hp["W"]) + hp["bias"]
    assert torch.allclose(out[0], ref, rtol=1e-12, atol=1e-12)


def test_residual_is_a_dead_store_when_widths_agree_and_a_conv1d_otherwise():
    """utils/layers.py:38-42."""
    cfg = synth.tiny(seed=23, n=30, f=4, p=1, deg=4.0)
    rng = np.random.default_rng(24)
    t = lambda *s: torch.from_numpy(rng.normal(size=s))
    X = torch.from_numpy(cfg.X).double()[None]
    bias = torch.from_numpy(O.adj_to_bias(cfg.adjs()[0], [cfg.N], 1))
    hp = {"W": t(4, 4), "a1": t(4), "b1": t(), "a2": t(4), "b2": t(), "bias": t(4)}
    assert torch.equal(O.attn_head(X, 4, bias, O.elu, hp, residual=True), O.attn_head(X, 4, bias, O.elu, hp))
    hp2 = {"W": t(4, 6), "a1": t(6), "b1": t(), "a2": t(6), "b2": t(), "bias": t(6), "W_res": t(4, 6), "b_res": t(6)}
    plain = O.attn_head(X, 6, bias, O.identity, hp2)
    res = O.attn_head(X, 6, bias, O.identity, hp2, residual=True)
    assert torch.allclose(res - plain, X @ hp2["W_res"] + hp2["b_res"], rtol=1e-12, atol=1e-12)


def test_stacked_layers_compose_attn_heads():
    """models/gat.py:48-57: layer 2 of meta-path p reads the concatenated heads of layer 1 of the SAME meta-path."""
    cfg = synth.tiny(seed=25, n=36, f=10, p=2, deg=5.0)
    params = O.init_params(np.random.default_rng(26), [cfg.F] * 2, cfg.C, hid=4, heads=2, mp_att_size=8, deep=[(3, 4)])
    X = torch.from_numpy(cfg.X).double()[None]
    biases = [torch.from_numpy(O.adj_to_bias(a, [cfg.N], 1)) for a in cfg.adjs()]
    _, fe, av = O.HeteGAT_multi_inference([X, X], cfg.C, cfg.N, False, 0.0, 0.0, biases, [4, 4], [2, 3, 1], params,
                                          mp_att_size=8)
    embeds = []
    for p in range(2):
        h1 = torch.cat([O.attn_head(X, 4, biases[p], O.elu, O.head_params(params, p, k)) for k in range(2)], -1)
        h2 = torch.cat([O.attn_head(h1, 4, biases[p], O.elu, O.head_params(params, p, k, layer=1)) for k in range(3)], -1)
        embeds.append(h2[0].unsqueeze(1))
    fe2, av2 = O.SimpleAttLayer(torch.cat(embeds, 1), 8, params, return_alphas=True)
    assert fe.shape == (cfg.N, 12) and torch.equal(fe, fe2) and torch.equal(av, av2)


def test_gat_inference_averages_identity_output_heads():
    """models/gat.py:25-30."""
    cfg = synth.tiny(seed=27, n=32, f=9, p=1, c=5, deg=5.0)
    params = O.init_gat_params(np.random.default_rng(28), cfg.F, cfg.C, [4], [2, 3])
    X = torch.from_numpy(cfg.X).double()[None]
    bias = torch.from_numpy(O.adj_to_bias(cfg.adjs()[0], [cfg.N], 1))
    logits = O.GAT_inference(X, cfg.C, cfg.N, False, 0.0, 0.0, bias, [4], [2, 3], params)
    lay, out = params["hidden"][0], params["out"]
    hd = lambda l, k, H: {"W": l["W"][:, k * H:(k + 1) * H], "a1": l["a1"][k], "b1": l["b1"][k], "a2": l["a2"][k],
                          "b2": l["b2"][k], "bias": l["bias"][k * H:(k + 1) * H]}
    h1 = torch.cat([O.attn_head(X, 4, bias, O.elu, hd(lay, k, 4)) for k in range(2)], -1)
    ref = sum(O.attn_head(h1, cfg.C, bias, O.identity, hd(out, k, cfg.C)) for k in range(3)) / 3
    assert logits.shape == (1, cfg.N, cfg.C) and torch.allclose(logits, ref, rtol=1e-12, atol=1e-12)
