"""K-A/K-B/K-D parity: projection + fused CSR edge-softmax-aggregate, forward and backward, against
the fp64 dense oracle (utils/layers.py:7-46 restated) on the same seeded inputs, through the same
autograd.Function -> ctypes -> C-ABI route the model uses."""
import numpy as np
import pytest
import torch

from han_b200 import synth
from oracle import han_oracle as O
from tests.util import assert_close, compare_step, oracle_step, product_step

pytestmark = pytest.mark.gpu


def _rand_params(rng, F, G, K, H):
    D = K * H
    t = lambda *s: torch.from_numpy(rng.normal(size=s) * 0.3)
    return {"W": t(F, G * D), "a1": t(G, K, H), "b1": t(G, K), "a2": t(G, K, H), "b2": t(G, K), "bias": t(G, D)}


def _oracle_node_attention(cfg, par, K, H, act, upstream):
    """fp64 dense oracle: Z (N,G,D) from G*K attn_head calls, and grads of sum(Z * upstream)."""
    G, D = cfg.P, K * H
    p = {k: v.clone().double().requires_grad_(True) for k, v in par.items()}
    X = torch.from_numpy(cfg.X).double().unsqueeze(0)
    biases = [torch.from_numpy(O.adj_to_bias(a, [cfg.N], 1)) for a in cfg.adjs()]
    cols, coefs = [], []
    for g in range(G):
        heads = []
        for k in range(K):
            hp = {"W": p["W"][:, g * D + k * H: g * D + (k + 1) * H], "a1": p["a1"][g, k], "b1": p["b1"][g, k],
                  "a2": p["a2"][g, k], "b2": p["b2"][g, k], "bias": p["bias"][g, k * H:(k + 1) * H]}
            o, c = O.attn_head(X, H, biases[g], act, hp, return_coef=True)
            heads.append(o[0]); coefs.append(c[0].detach())
        cols.append(torch.cat(heads, -1))
    Z = torch.stack(cols, 1)
    (Z * upstream.double()).sum().backward()
    return Z.detach(), {k: v.grad for k, v in p.items()}, coefs


def _product_node_attention(cfg, par, K, H, act_name, upstream, want_coefs=False):
    import han_b200 as hb
    from han_b200 import ops
    dev = torch.device("cuda")
    p = {k: v.float().to(dev).requires_grad_(True) for k, v in par.items()}
    X = torch.from_numpy(cfg.X).to(dev)
    graphs = [hb.process.adj_to_bias(a, [cfg.N]) for a in cfg.adjs()]
    plan = ops.NodeAttentionPlan(graphs=graphs, K=K, H=H, act=ops.activation_code(act_name), want_coefs=want_coefs)
    Z = ops.node_attention(plan, X, p["W"], p["a1"], p["b1"], p["a2"], p["b2"], p["bias"])
    (Z * upstream.to(dev)).sum().backward()
    torch.cuda.synchronize()
    return Z.detach(), {k: v.grad for k, v in p.items()}, plan


SHAPES = [(8, 8), (4, 8), (1, 8), (8, 4), (2, 8), (8, 16), (16, 4), (1, 4), (4, 16)]


def test_chunk_boundaries_on_a_larger_graph():
    """~75 chunks of 2048 edges: rows straddling batch and chunk boundaries, a row longer than a
    whole chunk, runs of empty transposed rows; checked against the fp64 edge-list twin (forward)
    and its autograd (backward)."""
    import han_b200 as hb
    from han_b200 import ops
    rng = np.random.default_rng(77)
    n, F, K, H = 5000, 24, 8, 8
    deg = rng.integers(1, 60, size=n)
    deg[1234] = 2400                                 # longer than one 2048-edge chunk
    rows = []
    for i in range(n):
        c = rng.choice(n // 2, size=deg[i], replace=False)   # sources only in the first half: the
        c[0] = min(i, n // 2 - 1) if i < n // 2 else c[0]     # transposed graph has 2500 empty rows
        rows.append(np.unique(c).astype(np.int32))
    indptr = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int64)
    indices = np.concatenate(rows)
    X = rng.normal(size=(n, F)).astype(np.float32)
    par = _rand_params(rng, F, 1, K, H)
    up = torch.from_numpy(rng.normal(size=(n, 1, K * H)))
    # oracle: edge-list twin per head
    p64 = {k: v.clone().double().requires_grad_(True) for k, v in par.items()}
    heads = []
    for k in range(K):
        hp = {"W": p64["W"][:, k * H:(k + 1) * H], "a1": p64["a1"][0, k], "b1": p64["b1"][0, k],
              "a2": p64["a2"][0, k], "b2": p64["b2"][0, k], "bias": p64["bias"][0, k * H:(k + 1) * H]}
        heads.append(O.attn_head_edges(torch.from_numpy(X).double(), indptr, indices, hp))
    Zo = torch.cat(heads, -1).unsqueeze(1)
    (Zo * up).sum().backward()
    dev = torch.device("cuda")
    p32 = {k: v.float().to(dev).requires_grad_(True) for k, v in par.items()}
    g = hb.MetaPathGraph.from_csr(indptr, indices, n_cols=n)
    plan = ops.NodeAttentionPlan(graphs=[g], K=K, H=H)
    Z = ops.node_attention(plan, torch.from_numpy(X).to(dev), p32["W"], p32["a1"], p32["b1"], p32["a2"], p32["b2"],
                           p32["bias"])
    (Z * up.float().to(dev)).sum().backward()
    torch.cuda.synchronize()
    assert_close(Z, Zo.detach(), "Z")
    for k in p64:
        assert_close(p32[k].grad, p64[k].grad, "d" + k)


@pytest.mark.parametrize("K,H", SHAPES)
def test_node_attention_fwd_bwd_parity(K, H):
    cfg = synth.tiny(seed=K * 31 + H, n=131, f=37, p=2, deg=7.0)   # N, F not multiples of 32/4
    rng = np.random.default_rng(K * 100 + H)
    par = _rand_params(rng, cfg.F, cfg.P, K, H)
    up = torch.from_numpy(rng.normal(size=(cfg.N, cfg.P, K * H)))
    Zo, go, _ = _oracle_node_attention(cfg, par, K, H, O.elu, up)
    Zp, gp, _ = _product_node_attention(cfg, par, K, H, "elu", up.float())
    assert_close(Zp, Zo, "Z")
    for k in go:
        assert_close(gp[k], go[k], "d" + k)


@pytest.mark.parametrize("mode,binary,rel", [(1, False, 1e-5), (2, True, 1e-5), (3, False, 1e-2)])
@pytest.mark.parametrize("n,F,P", [(300, 37, 2), (515, 256, 4), (130, 96, 5), (128, 32, 1)])
def test_tensor_core_projection(mode, binary, rel, n, F, P):
    """K-A on tcgen05 (TMA + TMEM): 3xTF32 must meet the FP32 bound, 2xTF32 too when X is exactly
    representable in tf32 (0/1 features), plain TF32 the looser stated bound."""
    import han_b200 as hb
    from han_b200 import ops
    cfg = synth.tiny(seed=n + F + P, n=n, f=F, p=P, deg=6.0, binary=binary)
    rng = np.random.default_rng(n * 7 + F)
    par = _rand_params(rng, F, P, 8, 8)
    up = torch.from_numpy(rng.normal(size=(n, P, 64)))
    Zo, go, _ = _oracle_node_attention(cfg, par, 8, 8, O.elu, up)
    dev = torch.device("cuda")
    p = {k: v.float().to(dev).requires_grad_(True) for k, v in par.items()}
    graphs = [hb.process.adj_to_bias(a, [cfg.N]) for a in cfg.adjs()]
    plan = ops.NodeAttentionPlan(graphs=graphs, K=8, H=8, project_mode=mode)
    Z = ops.node_attention(plan, torch.from_numpy(cfg.X).to(dev), p["W"], p["a1"], p["b1"], p["a2"], p["b2"], p["bias"])
    (Z * up.float().to(dev)).sum().backward()
    torch.cuda.synchronize()
    assert_close(Z, Zo, "Z", rel=rel, rtol=max(1e-4, 30 * rel))
    for k in go:
        # plain TF32 (mode 3): outputs within 1e-2, gradients (sums with cancellation) within 5e-2
        grel = rel if mode != 3 else 5e-2
        assert_close(p[k].grad, go[k], "d" + k, rel=grel, rtol=max(1e-4, 30 * grel))


def test_identity_activation_and_three_metapaths():
    cfg = synth.tiny(seed=5, n=64, f=16, p=3, deg=4.0)
    rng = np.random.default_rng(55)
    par = _rand_params(rng, cfg.F, cfg.P, 8, 8)
    up = torch.from_numpy(rng.normal(size=(cfg.N, cfg.P, 64)))
    Zo, go, _ = _oracle_node_attention(cfg, par, 8, 8, O.identity, up)
    Zp, gp, _ = _product_node_attention(cfg, par, 8, 8, "identity", up.float())
    assert_close(Zp, Zo, "Z")
    for k in go:
        assert_close(gp[k], go[k], "d" + k)


def test_degenerate_rows_and_long_rows():
    """single-neighbour rows (alpha == 1 exactly), a full row, a source every node attends to (long
    transposed row), rows longer than one 32-edge chunk, no-self-loop rows."""
    cfg = synth.tiny(seed=9, n=200, f=24, p=2, deg=3.0)
    m0, m1 = cfg.masks
    m0[3, :] = False; m0[3, 3] = True
    m0[7, :] = True
    m0[50:60, :150] = True
    m1[:, 11] = True
    m1[20, :] = False; m1[20, 5] = True
    rng = np.random.default_rng(99)
    par = _rand_params(rng, cfg.F, cfg.P, 8, 8)
    up = torch.from_numpy(rng.normal(size=(cfg.N, cfg.P, 64)))
    Zo, go, coefs_o = _oracle_node_attention(cfg, par, 8, 8, O.elu, up)
    Zp, gp, plan = _product_node_attention(cfg, par, 8, 8, "elu", up.float(), want_coefs=True)
    assert_close(Zp, Zo, "Z")
    for k in go:
        assert_close(gp[k], go[k], "d" + k)
    # return_coef: alpha on edges == dense coefs on edges; single-neighbour rows exactly 1
    for g in range(cfg.P):
        indptr, indices = plan.graphs[g].to_host()
        rows = np.repeat(np.arange(cfg.N), np.diff(indptr))
        alpha = plan.coefs[g].cpu()
        for k in range(8):
            ref = coefs_o[g * 8 + k][rows, indices]
            assert_close(alpha[:, k], ref, f"alpha[{g}][{k}]")
        rowsum = torch.zeros(cfg.N, 8).index_add_(0, torch.from_numpy(rows), alpha)
        assert torch.allclose(rowsum, torch.ones_like(rowsum), atol=2e-6)     # softmax rows sum to 1
    a0 = plan.coefs[0].cpu()
    ip0 = plan.graphs[0].to_host()[0]
    assert torch.equal(a0[ip0[3]], torch.ones(8))                                # alpha == 1 exactly


def test_row_without_any_edge_matches_fp32_dense_forward():
    """an all -1e9 bias row degenerates to uniform attention over ALL nodes in the
    reference's fp32 arithmetic; compare with the fp32 oracle (fp64 does not absorb the logits)."""
    import han_b200 as hb
    from han_b200 import ops
    cfg = synth.tiny(seed=12, n=90, f=10, p=1, deg=4.0)
    cfg.masks[0][17, :] = False          # with adj = mask - I this row has adj_ii = -1 -> no self-loop
    rng = np.random.default_rng(3)
    par = _rand_params(rng, cfg.F, 1, 8, 8)
    bias32 = torch.from_numpy(O.adj_to_bias(cfg.adjs()[0], [cfg.N], 1)).float()
    X32 = torch.from_numpy(cfg.X).unsqueeze(0)
    heads = []
    for k in range(8):
        hp = {"W": par["W"][:, k * 8:(k + 1) * 8].float(), "a1": par["a1"][0, k].float(), "b1": par["b1"][0, k].float(),
              "a2": par["a2"][0, k].float(), "b2": par["b2"][0, k].float(), "bias": par["bias"][0, k * 8:(k + 1) * 8].float()}
        heads.append(O.attn_head(X32, 8, bias32, O.elu, hp)[0])
    Zo = torch.cat(heads, -1)
    dev = torch.device("cuda")
    g = hb.process.adj_to_bias(cfg.adjs()[0], [cfg.N])
    assert g.has_empty_rows()
    plan = ops.NodeAttentionPlan(graphs=[g], K=8, H=8)
    p = {k: v.float().to(dev) for k, v in par.items()}
    with torch.no_grad():
        Z = ops.node_attention(plan, torch.from_numpy(cfg.X).to(dev), p["W"], p["a1"], p["b1"], p["a2"], p["b2"], p["bias"])
    assert_close(Z[:, 0], Zo, "Z", rel=5e-6)


def test_attn_head_reference_signature():
    """layers.attn_head keeps the reference's positional signature (utils/layers.py:7-8)."""
    import han_b200 as hb
    cfg = synth.tiny(seed=2, n=77, f=19, p=1, deg=5.0)
    rng = np.random.default_rng(8)
    H = 8
    hp64 = {"W": torch.from_numpy(rng.normal(size=(cfg.F, H)) * 0.3), "a1": torch.from_numpy(rng.normal(size=H)),
            "b1": torch.tensor(0.1, dtype=torch.float64), "a2": torch.from_numpy(rng.normal(size=H)),
            "b2": torch.tensor(-0.3, dtype=torch.float64), "bias": torch.from_numpy(rng.normal(size=H))}
    bias = O.adj_to_bias(cfg.adjs()[0], [cfg.N], 1)
    ref, coefs = O.attn_head(torch.from_numpy(cfg.X).double()[None], H, torch.from_numpy(bias), O.elu, hp64,
                             return_coef=True)
    dev = torch.device("cuda")
    hp = {k: v.float().to(dev) for k, v in hp64.items()}
    seq = torch.from_numpy(cfg.X).to(dev)[None]
    # dense fp32 bias, exactly what the reference driver feeds (ex_acm3025.py:127,180)
    out, ec = hb.layers.attn_head(seq, H, torch.from_numpy(bias.astype(np.float32)), hb.layers.elu, 0.0, 0.0, False,
                                  True, params=hp)
    assert out.shape == (1, cfg.N, H)
    assert_close(out, ref, "attn_head")
    assert_close(ec.to_dense(), coefs, "coefs")
    assert (ec.to_dense()[0].cpu()[torch.from_numpy(bias[0]) < 0] == 0).all()    # masked coefficients exactly 0
    # the reference's training-time call shape (in_drop = coef_drop = 0.6) runs and differs from eval
    dropped = hb.layers.attn_head(seq, H, torch.from_numpy(bias.astype(np.float32)), hb.layers.elu, 0.6, 0.6, params=hp)
    assert dropped.shape == out.shape and torch.isfinite(dropped).all() and not torch.equal(dropped, out)


@pytest.mark.parametrize("split", [16, 50])
def test_heavy_rows_cut_into_segments_match_oracle(split, monkeypatch):
    """Power-law meta-paths (BASELINE.json configs[4]): hub rows AND hub columns far longer than the
    split length are processed as several virtual rows + a merge (MetaPathGraph.split_view,
    han_attn_*_chunked_split); the whole step must still match the dense fp64 oracle, and equal the
    un-split kernels up to summation order."""
    import han_b200 as hb
    from han_b200 import graph as hg
    cfg = synth.tiny(seed=141, n=260, f=20, p=2, deg=5.0)
    rng = np.random.default_rng(142)
    for m in cfg.masks:                       # hubs: 3 nodes linked to (almost) everyone, both directions
        hubs = rng.choice(cfg.N, size=3, replace=False)
        for h in hubs:
            sel = rng.random(cfg.N) < 0.9
            m[h, sel] = True
            m[sel, h] = True
        m[rng.integers(0, cfg.N), :] = True   # one row that is exactly full
    params = O.init_params(np.random.default_rng(143), [cfg.F] * cfg.P, cfg.C)
    out_o, grads_o = oracle_step(cfg, params)
    out_u, grads_u, _ = product_step(cfg, params)                  # default split length: nothing is cut
    monkeypatch.setattr(hg, "SPLIT_ROW_EDGES", split)
    graphs = [hb.process.adj_to_bias(a, [cfg.N]) for a in cfg.adjs()]
    sv, tv = graphs[0].split_view(), graphs[0].transpose().split_view()
    assert sv is not None and tv is not None and sv.n_heavy >= 4 and tv.n_heavy >= 3
    indptr = graphs[0].indptr.cpu().numpy()
    iv = sv.indptr_v.cpu().numpy()
    assert set(indptr) <= set(iv) and np.diff(iv).max() <= split and iv[-1] == graphs[0].nnz
    vm = sv.vmap.cpu().numpy()
    assert (np.bincount(vm[:, 0], minlength=cfg.N) == np.maximum(1, -(-np.diff(indptr) // split))).all()
    assert sv.n_slots == int((vm[:, 1] >= 0).sum()) == int(sv.heavy_ptr[-1])
    out_p, grads_p, _ = product_step(cfg, params, graphs=graphs)
    compare_step(out_o, grads_o, out_p, grads_p)
    assert torch.allclose(out_p["logits"], out_u["logits"], rtol=1e-5, atol=1e-6)
    assert torch.allclose(grads_p["W"][0], grads_u["W"][0], rtol=1e-4, atol=1e-7)


def test_df1_is_row_local_on_plain_and_heavy_row_graphs(monkeypatch):
    """The forward keeps a second aggregate (V' = sum alpha~ leaky' S, c = sum alpha leaky'), so that
    df1_i = <dV_i, V'_i> - delta_i c_i needs no per-edge array and no by-destination pass; same gradients on a graph
    with a full row / a hub column, un-split and through the virtual-row (split + merge) kernels."""
    from han_b200 import graph as hg
    cfg = synth.tiny(seed=161, n=230, f=18, p=2, deg=7.0)
    cfg.masks[0][5, :] = True
    cfg.masks[0][:, 9] = True
    params = O.init_params(np.random.default_rng(162), [cfg.F] * cfg.P, cfg.C)
    out_o, grads_o = oracle_step(cfg, params)
    out_p, grads_p, _ = product_step(cfg, params)
    compare_step(out_o, grads_o, out_p, grads_p)
    monkeypatch.setattr(hg, "SPLIT_ROW_EDGES", 32)          # and through the virtual-row kernels
    out_s, grads_s, _ = product_step(cfg, params)
    compare_step(out_o, grads_o, out_s, grads_s)


@pytest.mark.parametrize("cfg", ["1,1", "2,2", "3,3"])
def test_gather_ring_geometries(cfg):
    """HAN_GATHER_CFG (attn_stream.cu): the alternative ring geometries kept for tuning runs (tools/gather_sweep.sh) are
    the same kernels with other template parameters; each must pass the same step parity as the default."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, HAN_GATHER_CFG=cfg, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "gather_cfg_check.py")], env=env, cwd=root,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "gather cfg ok" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]

