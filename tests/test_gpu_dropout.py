"""Training-mode dropout (utils/layers.py:18-19,29-32; ex_acm3025.py:185-186 feeds 0.6/0.6).

TF's random stream cannot be reproduced, so the check is: the kernels' masks are pure functions of
(seed, meta-path, head, coordinates) -- replicated here in numpy -- and with those SAME masks handed to
the oracle (its `masks=` hook applies them at the three tf.nn.dropout sites) the CUDA forward and every
gradient must match to the usual 1e-5.  Plus rate / independence statistics of the masks themselves."""
import numpy as np
import pytest
import torch

from han_b200 import synth
from oracle import han_oracle as O
from tests.util import assert_close, assert_head_grads_close

M32 = np.uint64(0xFFFFFFFF)


def mix3(seed, a, b):
    """numpy replica of han_rng.cuh::mix3 (uint32 wrap-around arithmetic done in uint64)."""
    seed, a, b = (np.asarray(v, dtype=np.uint64) for v in (seed, a, b))
    h = (seed ^ np.uint64(0x9E3779B9)) & M32
    h = ((h ^ a) * np.uint64(0x85EBCA6B)) & M32
    h ^= h >> np.uint64(13)
    h = ((h ^ b) * np.uint64(0xC2B2AE35)) & M32
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x27D4EB2F)) & M32
    h ^= h >> np.uint64(15)
    return h


def stream_seed(seed, purpose, metapath, head):
    return mix3(seed, (purpose * 0x01000193 + metapath) & 0xFFFFFFFF, (head + 0x7F4A7C15) & 0xFFFFFFFF)


def thr_of(keep):
    return np.uint64(int(np.float32(keep) * np.float32(16777216.0) + np.float32(0.5)))


def coef_mask(seed, g, k, n_dst, n_src, keep):
    i, j = np.meshgrid(np.arange(n_dst), np.arange(n_src), indexing="ij")
    return (mix3(stream_seed(seed, 3, g, k), i, j) >> np.uint64(8)) < thr_of(keep)


def s_mask(seed, g, n, D, keep):
    i, d = np.meshgrid(np.arange(n), np.arange(D), indexing="ij")
    return (mix3(stream_seed(seed, 2, g, 0), i, d) >> np.uint64(8)) < thr_of(keep)


def x_mask(seed, g, k, n, F, keep):
    i, f = np.meshgrid(np.arange(n), np.arange(F), indexing="ij")
    base = mix3(stream_seed(seed, 1, g, 0), i, f)
    h = (((base ^ np.uint64((0x632BE5AB * (k + 1)) & 0xFFFFFFFF)) * np.uint64(0x9E3779B1)) & M32) >> np.uint64(8)
    return h < thr_of(keep)


def test_mask_statistics():
    """keep rate, independence across heads / meta-paths / seeds (CPU: the replica is what the GPU
    parity test below ties to the kernels)."""
    for keep in (0.4, 0.9):
        m = coef_mask(12345, 0, 3, 400, 400, keep)
        assert abs(m.mean() - keep) < 4 * np.sqrt(keep * (1 - keep) / m.size)
        x0, x1 = x_mask(777, 1, 0, 300, 200, keep), x_mask(777, 1, 5, 300, 200, keep)
        assert abs(x0.mean() - keep) < 0.01 and abs(x1.mean() - keep) < 0.01
        both = (x0 & x1).mean()
        assert abs(both - keep * keep) < 0.01                      # heads draw independent masks
        s0, s1 = s_mask(1, 0, 500, 64, keep), s_mask(2, 0, 500, 64, keep)
        assert abs((s0 & s1).mean() - keep * keep) < 0.01           # seeds decorrelate


@pytest.mark.gpu
@pytest.mark.parametrize("in_drop,coef_drop", [(0.6, 0.6), (0.0, 0.5), (0.3, 0.0)])
def test_dropout_matches_oracle_with_same_masks(in_drop, coef_drop):
    import han_b200 as hb
    from han_b200 import ops
    K, H, D = 8, 8, 64
    cfg = synth.tiny(seed=77, n=90, f=28, p=2, deg=7.0)
    rng = np.random.default_rng(78)
    t = lambda *s: torch.from_numpy(rng.normal(size=s) * 0.3)
    G, F, n = cfg.P, cfg.F, cfg.N
    par = {"W": t(F, G * D), "a1": t(G, K, H), "b1": t(G, K), "a2": t(G, K, H), "b2": t(G, K), "bias": t(G, D)}
    up = torch.from_numpy(rng.normal(size=(n, G, D)))
    dev = torch.device("cuda")
    seed = torch.tensor([20251018], dtype=torch.int32, device=dev)
    p = {k: v.float().to(dev).requires_grad_(True) for k, v in par.items()}
    graphs = [hb.process.adj_to_bias(a, [n]) for a in cfg.adjs()]
    plan = ops.NodeAttentionPlan(graphs=graphs, K=K, H=H, in_drop=in_drop, coef_drop=coef_drop, seed=seed,
                                 metapath_ids=[0, 1])
    Z = ops.node_attention(plan, torch.from_numpy(cfg.X).to(dev), p["W"], p["a1"], p["b1"], p["a2"], p["b2"], p["bias"])
    (Z * up.float().to(dev)).sum().backward()
    torch.cuda.synchronize()

    sv = int(seed.item()) & 0xFFFFFFFF
    p64 = {k: v.clone().double().requires_grad_(True) for k, v in par.items()}
    X = torch.from_numpy(cfg.X).double()[None]
    biases = [torch.from_numpy(O.adj_to_bias(a, [n], 1)) for a in cfg.adjs()]
    cols = []
    for g in range(G):
        sm = s_mask(sv, g, n, D, 1.0 - in_drop) if in_drop else np.ones((n, D), bool)
        heads = []
        for k in range(K):
            hp = {"W": p64["W"][:, g * D + k * H:g * D + (k + 1) * H], "a1": p64["a1"][g, k], "b1": p64["b1"][g, k],
                  "a2": p64["a2"][g, k], "b2": p64["b2"][g, k], "bias": p64["bias"][g, k * H:(k + 1) * H]}
            masks = {"x": torch.from_numpy(x_mask(sv, g, k, n, F, 1.0 - in_drop) if in_drop else np.ones((n, F), bool)),
                     "coef": torch.from_numpy(coef_mask(sv, g, k, n, n, 1.0 - coef_drop) if coef_drop
                                              else np.ones((n, n), bool)),
                     "s": torch.from_numpy(sm[:, k * H:(k + 1) * H])}
            heads.append(O.attn_head(X, H, biases[g], O.elu, hp, in_drop=in_drop, coef_drop=coef_drop, masks=masks)[0])
        cols.append(torch.cat(heads, -1))
    Zo = torch.stack(cols, 1)
    (Zo * up).sum().backward()
    assert_close(Z, Zo.detach(), "Z")
    for k in p64:
        assert_close(p[k].grad, p64[k].grad, "d" + k)


@pytest.mark.gpu
def test_reference_training_call_with_dropout():
    """The reference driver's training feed (attn_drop = ffd_drop = 0.6, ex_acm3025.py:185-186) runs
    through the drop-in API; masks change from call to call; evaluation (0.0) is deterministic; the
    dropped forward is unbiased (mean over many calls approaches the un-dropped embedding scale)."""
    import han_b200 as hb
    cfg = synth.tiny(seed=5, n=200, f=30, p=2, deg=8.0, binary=True)
    dev = torch.device("cuda")
    hp = hb.HANParams([cfg.F] * 2, cfg.C, device=dev, generator=torch.Generator().manual_seed(3))
    X = torch.from_numpy(cfg.X).to(dev)[None]
    graphs = [hb.process.adj_to_bias(a, [cfg.N]) for a in cfg.adjs()]
    labels = torch.from_numpy(cfg.labels).to(dev)
    mask = torch.from_numpy(cfg.train_mask.astype(np.float32)).to(dev)
    train = hb.BaseGAttN.training(hp, 0.005, 0.001)
    outs = []
    for _ in range(3):
        logits, emb, att = hb.HeteGAT_multi.inference([X, X], cfg.C, cfg.N, True, 0.6, 0.6, graphs, [8], [8, 1], params=hp)
        loss = hb.BaseGAttN.masked_softmax_cross_entropy(logits.reshape(-1, cfg.C), labels, mask)
        train.opt.zero_grad()
        (loss + train.l2_loss()).backward()
        assert all(torch.isfinite(q.grad).all() for q in hp.parameters())
        outs.append(emb.detach().clone())
    assert not torch.equal(outs[0], outs[1])                     # fresh masks every call
    with torch.no_grad():
        e0 = hb.HeteGAT_multi.inference([X, X], cfg.C, cfg.N, False, 0.0, 0.0, graphs, [8], [8, 1], params=hp)[1]
        e1 = hb.HeteGAT_multi.inference([X, X], cfg.C, cfg.N, False, 0.0, 0.0, graphs, [8], [8, 1], params=hp)[1]
    assert torch.equal(e0, e1)


def _head_oracle_with_product_masks(cfg, hp64, seed_value, g, in_drop, coef_drop, residual):
    n, F, H = cfg.N, cfg.F, hp64["W"].shape[1]
    masks = {"x": torch.from_numpy(x_mask(seed_value, g, 0, n, F, 1.0 - in_drop)),
             "coef": torch.from_numpy(coef_mask(seed_value, g, 0, n, n, 1.0 - coef_drop)),
             "s": torch.from_numpy(s_mask(seed_value, g, n, H, 1.0 - in_drop))}
    X = torch.from_numpy(cfg.X).double()[None]
    bias = torch.from_numpy(O.adj_to_bias(cfg.adjs()[0], [n], 1))
    return O.attn_head(X, H, bias, O.elu, hp64, in_drop=in_drop, coef_drop=coef_drop, residual=residual, masks=masks)


@pytest.mark.gpu
def test_residual_conv_reads_the_dropped_input():
    """residual=True in training mode (the reference's training configuration with residual, ex_acm3025.py:185-186):
    the residual conv1d of utils/layers.py:40 reads `seq` AFTER :19 reassigned it to the dropped copy.  Pinned to the
    reference by tests/golden/ref_attn_head_dropout_residual.npz (oracle == reference given the same masks); here the
    CUDA path == oracle given the product's masks, forward and every gradient."""
    import han_b200 as hb
    from han_b200 import layers
    cfg = synth.tiny(seed=301, n=80, f=20, p=1, deg=6.0)
    rng = np.random.default_rng(302)
    H, F = 8, cfg.F
    hp = {"W": rng.normal(size=(F, H)) * 0.3, "a1": rng.normal(size=H), "b1": np.float64(0.05), "a2": rng.normal(size=H),
          "b2": np.float64(-0.03), "bias": rng.normal(0, 0.1, H), "W_res": rng.normal(size=(F, H)) * 0.3,
          "b_res": rng.normal(0, 0.1, H)}
    dev = torch.device("cuda", torch.cuda.current_device())
    layers._DROP_SEEDS[str(dev)] = torch.tensor([4242], dtype=torch.int32, device=dev)      # attn_head bumps it to 4243
    pp = {k: torch.nn.Parameter(torch.as_tensor(v).float().to(dev)) for k, v in hp.items()}
    graph = hb.process.adj_to_bias(cfg.adjs()[0], [cfg.N])
    up = torch.from_numpy(rng.normal(size=(1, cfg.N, H)))
    out = hb.layers.attn_head(torch.from_numpy(cfg.X).to(dev)[None], H, graph, hb.layers.elu, in_drop=0.6, coef_drop=0.6,
                              residual=True, params=pp)
    (out * up.float().to(dev)).sum().backward()
    p64 = {k: torch.as_tensor(v).double().clone().requires_grad_(True) for k, v in hp.items()}
    ref = _head_oracle_with_product_masks(cfg, p64, 4243, 0, 0.6, 0.6, True)
    (ref * up).sum().backward()
    assert_close(out, ref.detach(), "out")
    assert_head_grads_close({k: v.grad for k, v in pp.items()}, {k: v.grad for k, v in p64.items()})


@pytest.mark.gpu
def test_several_dropped_calls_then_one_backward_keep_their_own_masks():
    """The reference's own pattern (models/gat.py:42-46): attn_head is called K times, THEN the graph is
    differentiated.  Every call must regenerate ITS masks in the backward (a per-call snapshot of the seed word),
    not those of the last call."""
    import han_b200 as hb
    from han_b200 import layers
    cfg = synth.tiny(seed=311, n=70, f=16, p=1, deg=6.0)
    rng = np.random.default_rng(312)
    H, F = 8, cfg.F
    dev = torch.device("cuda", torch.cuda.current_device())
    layers._DROP_SEEDS[str(dev)] = torch.tensor([900], dtype=torch.int32, device=dev)
    graph = hb.process.adj_to_bias(cfg.adjs()[0], [cfg.N])
    X = torch.from_numpy(cfg.X).to(dev)[None]
    hps, pps, outs = [], [], []
    for c in range(3):
        hp = {"W": rng.normal(size=(F, H)) * 0.3, "a1": rng.normal(size=H), "b1": np.float64(0.0), "a2": rng.normal(size=H),
              "b2": np.float64(0.0), "bias": rng.normal(0, 0.1, H)}
        pp = {k: torch.nn.Parameter(torch.as_tensor(v).float().to(dev)) for k, v in hp.items()}
        outs.append(hb.layers.attn_head(X, H, graph, hb.layers.elu, in_drop=0.5, coef_drop=0.5, params=pp))
        hps.append(hp); pps.append(pp)
    up = torch.from_numpy(rng.normal(size=(1, cfg.N, 3 * H)))
    (torch.cat(outs, -1) * up.float().to(dev)).sum().backward()          # ONE backward after all forwards
    for c in range(3):
        p64 = {k: torch.as_tensor(v).double().clone().requires_grad_(True) for k, v in hps[c].items()}
        ref = _head_oracle_with_product_masks(cfg, p64, 901 + c, 0, 0.5, 0.5, False)
        (ref * up[..., c * H:(c + 1) * H]).sum().backward()
        assert_close(outs[c], ref.detach(), f"out[{c}]")
        for k in p64:
            assert_close(pps[c][k].grad, p64[k].grad, f"call {c}: d{k}")


@pytest.mark.gpu
def test_stacked_residual_model_trains_with_dropout():
    """HeteGAT_multi with hid_units=[8,8], residual=True and the reference's 0.6 / 0.6 dropout feed: runs, gives finite
    gradients for every variable (incl. W_res / b_res), fresh masks per call."""
    import han_b200 as hb
    cfg = synth.tiny(seed=321, n=150, f=24, p=2, deg=7.0, binary=True)
    dev = torch.device("cuda")
    hp = hb.HANParams([cfg.F] * 2, cfg.C, (8, 8), (4, 2, 1), 32, device=dev, residual=True,
                      generator=torch.Generator().manual_seed(5))
    X = torch.from_numpy(cfg.X).to(dev)[None]
    graphs = [hb.process.adj_to_bias(a, [cfg.N]) for a in cfg.adjs()]
    labels = torch.from_numpy(cfg.labels).to(dev)
    mask = torch.from_numpy(cfg.train_mask.astype(np.float32)).to(dev)
    embs = []
    for _ in range(2):
        logits, emb, _ = hb.HeteGAT_multi.inference([X, X], cfg.C, cfg.N, True, 0.6, 0.6, graphs, [8, 8], [4, 2, 1],
                                                    residual=True, mp_att_size=32, params=hp)
        hb.BaseGAttN.masked_softmax_cross_entropy(logits.reshape(-1, cfg.C), labels, mask).backward()
        embs.append(emb.detach().clone())
    assert all(q.grad is not None and torch.isfinite(q.grad).all() for q in hp.parameters())
    assert float(hp.deep[0]["W_res"][0].grad.abs().max()) > 0
    assert not torch.equal(embs[0], embs[1])
