"""Multi-rank parity check, run under torchrun (one process per GPU):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_check.py

Every rank builds the same seeded inputs, takes its destination-row shard, runs the sharded
fwd+bwd through HeteGAT_multi.inference(dist=...) and compares its rows of the outputs, the
all-reduced loss and the all-reduced gradients with the fp64 dense oracle of the WHOLE graph."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import han_b200 as hb
from han_b200 import dist as hd, synth
from oracle import han_oracle as O
from tests.util import assert_close, oracle_step


def run_case(shard, cfg, params, mode):
    dev = shard.device
    N, P, C = cfg.N, cfg.P, cfg.C
    lo, hi = shard.row_range(N)
    out_o, grads_o = oracle_step(cfg, params, semantic_mode=mode)
    hp = hb.HANParams([cfg.F] * P, C, device=dev).load_dict(params)
    full = [hb.process.adj_to_bias(a, [N]) for a in cfg.adjs()]
    graphs = [g.row_slice(lo, hi) for g in full]
    shard._bwd = {}
    shard.bind(graphs, N)
    X = torch.from_numpy(cfg.X[lo:hi]).to(dev)[None]
    labels = torch.from_numpy(cfg.labels[lo:hi]).to(dev)
    mask = torch.from_numpy(cfg.train_mask[lo:hi].astype(np.float32)).to(dev)
    train = hb.BaseGAttN.training(hp, 0.005, 0.001)
    logits, fe, av = hb.HeteGAT_multi.inference([X] * P, C, N, True, 0.0, 0.0, graphs, [8], [8, 1], params=hp,
                                                semantic_mode=mode, dist=shard)
    total = shard.masked_loss(logits.reshape(-1, C), labels, mask, train)
    total.backward()
    shard.all_reduce_grads(hp)
    tot = shard.all_reduce_sum(total.detach().clone().reshape(1))
    torch.cuda.synchronize()
    assert_close(logits[0], out_o["logits"][0, lo:hi], "logits shard")
    assert_close(fe, out_o["final_embed"][lo:hi], "final_embed shard")
    assert_close(av, out_o["att_val"][lo:hi], "att_val shard")
    assert_close(tot[0], out_o["total"], "loss")
    gp = hp.grad_dict()
    for k, v in grads_o.items():
        if isinstance(v, list):
            for i, g in enumerate(v):
                assert_close(gp[k][i], g, f"d{k}[{i}]")
        else:
            assert_close(gp[k], v, f"d{k}")


def run_stacked_case(shard):
    """Several plans alive between forward and backward: stacked layers (hid_units=[8,8], one plan per meta-path
    and layer) AND a different feature tensor per meta-path (one group each).  Every plan invocation must own its
    symmetric-memory table set (the backward reads T / R from the tables)."""
    cfg = synth.tiny(seed=81, n=203, f=24, p=2, deg=6.0)
    rng = np.random.default_rng(82)
    params = O.init_params(rng, [cfg.F] * cfg.P, cfg.C, hid=8, heads=4, mp_att_size=32, deep=[(2, 8)], residual=True)
    X2 = rng.normal(size=cfg.X.shape).astype(np.float32)
    N, P, C = cfg.N, cfg.P, cfg.C
    dev = shard.device
    lo, hi = shard.row_range(N)
    p64 = O.params_to(params, torch.float64, requires_grad=True)
    Xs = [torch.from_numpy(cfg.X).double()[None], torch.from_numpy(X2).double()[None]]
    biases = [torch.from_numpy(O.adj_to_bias(a, [N], 1)) for a in cfg.adjs()]
    labels64 = torch.from_numpy(cfg.labels).double()
    mask64 = torch.from_numpy(cfg.train_mask.astype(np.float64))
    total_o, _, logits_o, fe_o, _ = O.step_loss(Xs, biases, labels64, mask64, p64, C, [8, 8], [4, 2, 1], 0.001, "reference",
                                               residual=True)
    total_o.backward()
    hp = hb.HANParams([cfg.F] * P, C, (8, 8), (4, 2, 1), 32, device=dev, residual=True).load_dict(params)
    full = [hb.process.adj_to_bias(a, [N]) for a in cfg.adjs()]
    graphs = [g.row_slice(lo, hi) for g in full]
    shard._bwd = {}
    shard.bind(graphs, N)
    Xl = [torch.from_numpy(cfg.X[lo:hi]).to(dev)[None], torch.from_numpy(X2[lo:hi]).to(dev)[None]]
    labels = torch.from_numpy(cfg.labels[lo:hi]).to(dev)
    mask = torch.from_numpy(cfg.train_mask[lo:hi].astype(np.float32)).to(dev)
    train = hb.BaseGAttN.training(hp, 0.005, 0.001)
    logits, fe, _ = hb.HeteGAT_multi.inference(Xl, C, N, True, 0.0, 0.0, graphs, [8, 8], [4, 2, 1], residual=True,
                                               mp_att_size=32, params=hp, dist=shard)
    total = shard.masked_loss(logits.reshape(-1, C), labels, mask, train)
    total.backward()
    shard.all_reduce_grads(hp)
    tot = shard.all_reduce_sum(total.detach().clone().reshape(1))
    torch.cuda.synchronize()
    assert_close(logits[0], logits_o.detach()[0, lo:hi], "stacked: logits shard")
    assert_close(fe, fe_o.detach()[lo:hi], "stacked: final_embed shard")
    assert_close(tot[0], total_o.detach(), "stacked: loss")
    gp = hp.grad_dict()
    for k in ("W", "a1", "a2", "bias"):
        for i in range(P):
            assert_close(gp[k][i], p64[k][i].grad, f"stacked: d{k}[{i}]")
    for kk, vv in p64["deep"][0].items():
        for i, t in enumerate(vv):
            assert_close(gp["deep"][0][kk][i], t.grad, f"stacked: deep.d{kk}[{i}]")


def run_dropout_case(shard):
    """Sharded node attention with training-mode dropout: masks are keyed by GLOBAL node ids, so the
    sharded run must equal the whole-graph oracle fed the same masks (numpy replica of han_rng.cuh)."""
    import torch.distributed as td
    from han_b200 import ops
    from tests.test_gpu_dropout import coef_mask, s_mask, x_mask
    K, H, D, in_drop, coef_drop = 8, 8, 64, 0.5, 0.4
    cfg = synth.tiny(seed=91, n=150, f=20, p=2, deg=6.0)
    rng = np.random.default_rng(92)
    t = lambda *s: torch.from_numpy(rng.normal(size=s) * 0.3)
    G, F, n = cfg.P, cfg.F, cfg.N
    par = {"W": t(F, G * D), "a1": t(G, K, H), "b1": t(G, K), "a2": t(G, K, H), "b2": t(G, K), "bias": t(G, D)}
    up = torch.from_numpy(rng.normal(size=(n, G, D)))
    dev = shard.device
    lo, hi = shard.row_range(n)
    seed = torch.tensor([424242], dtype=torch.int32, device=dev)
    p = {k: v.float().to(dev).requires_grad_(True) for k, v in par.items()}
    full = [hb.process.adj_to_bias(a, [n]) for a in cfg.adjs()]
    graphs = [g.row_slice(lo, hi) for g in full]
    shard._bwd = {}
    shard.bind(graphs, n)
    plan = ops.NodeAttentionPlan(graphs=graphs, K=K, H=H, in_drop=in_drop, coef_drop=coef_drop, seed=seed,
                                 metapath_ids=[0, 1], dist=shard)
    Z = ops.node_attention(plan, torch.from_numpy(cfg.X[lo:hi]).to(dev), p["W"], p["a1"], p["b1"], p["a2"], p["b2"],
                           p["bias"])
    (Z * up[lo:hi].float().to(dev)).sum().backward()
    for v in p.values():
        td.all_reduce(v.grad)
    torch.cuda.synchronize()
    sv = 424242
    p64 = {k: v.clone().double().requires_grad_(True) for k, v in par.items()}
    X = torch.from_numpy(cfg.X).double()[None]
    biases = [torch.from_numpy(O.adj_to_bias(a, [n], 1)) for a in cfg.adjs()]
    cols = []
    for g in range(G):
        sm = s_mask(sv, g, n, D, 1.0 - in_drop)
        heads = []
        for k in range(K):
            hp = {"W": p64["W"][:, g * D + k * H:g * D + (k + 1) * H], "a1": p64["a1"][g, k], "b1": p64["b1"][g, k],
                  "a2": p64["a2"][g, k], "b2": p64["b2"][g, k], "bias": p64["bias"][g, k * H:(k + 1) * H]}
            masks = {"x": torch.from_numpy(x_mask(sv, g, k, n, F, 1.0 - in_drop)),
                     "coef": torch.from_numpy(coef_mask(sv, g, k, n, n, 1.0 - coef_drop)),
                     "s": torch.from_numpy(sm[:, k * H:(k + 1) * H])}
            heads.append(O.attn_head(X, H, biases[g], O.elu, hp, in_drop=in_drop, coef_drop=coef_drop, masks=masks)[0])
        cols.append(torch.cat(heads, -1))
    Zo = torch.stack(cols, 1)
    (Zo * up).sum().backward()
    assert_close(Z, Zo.detach()[lo:hi], "Z shard (dropout)")
    for k in p64:
        assert_close(p[k].grad, p64[k].grad, "d" + k + " (dropout)")


def run_tile_case(tile, cfg, params, mode):
    """(meta-path x row-block) tiles (han_b200/tiles.py): this rank's meta-paths on its attention rows, the
    all-to-all of Z, semantic layer and loss on its semantic rows -- against the whole-graph oracle."""
    dev = tile.device
    N, P, C = cfg.N, cfg.P, cfg.C
    (a_lo, a_hi), (s_lo, s_hi) = tile.rows(N)
    out_o, grads_o = oracle_step(cfg, params, semantic_mode=mode)
    hp = hb.HANParams([cfg.F] * P, C, device=dev).load_dict(params)
    full = [hb.process.adj_to_bias(a, [N]) for a in cfg.adjs()]
    graphs = [full[p].row_slice(a_lo, a_hi) for p in tile.paths]
    tile.reset()
    tile.bind(graphs, N)
    X = torch.from_numpy(cfg.X[a_lo:a_hi]).to(dev)[None]
    labels = torch.from_numpy(cfg.labels[s_lo:s_hi]).to(dev)
    mask = torch.from_numpy(cfg.train_mask[s_lo:s_hi].astype(np.float32)).to(dev)
    train = hb.BaseGAttN.training(hp, 0.005, 0.001)
    logits, fe, av = hb.HeteGAT_multi.inference([X] * len(tile.paths), C, N, True, 0.0, 0.0, graphs, [8], [8, 1],
                                                params=hp, semantic_mode=mode, dist=tile)
    total = tile.masked_loss(logits.reshape(-1, C), labels, mask, train)
    total.backward()
    tile.all_reduce_grads(hp)
    tot = tile.all_reduce_sum(total.detach().clone().reshape(1))
    torch.cuda.synchronize()
    if s_hi > s_lo:
        assert_close(logits[0], out_o["logits"][0, s_lo:s_hi], "logits tile")
        assert_close(fe, out_o["final_embed"][s_lo:s_hi], "final_embed tile")
        assert_close(av, out_o["att_val"][s_lo:s_hi], "att_val tile")
    assert_close(tot[0], out_o["total"], "loss")
    gp = hp.grad_dict()
    for k, v in grads_o.items():
        if isinstance(v, list):
            for i, g in enumerate(v):
                assert_close(gp[k][i], g, f"d{k}[{i}]")
        else:
            assert_close(gp[k], v, f"d{k}")


def main_tile():
    from han_b200.tiles import TileShard
    base = hd.RowShard.init_process_group()
    W = base.world
    for seed, n, mode, P in ((71, 257, "reference", 4), (72, 310, "paper", 4), (73, 96, "reference", 2 if W <= 2 else W // 2)):
        if (W >= P and W % P) or (W < P and P % W):
            continue
        tile = TileShard(base.rank, W, P, base.device)
        cfg = synth.tiny(seed=seed, n=n, f=36, p=P, deg=6.0)
        cfg.masks[1][:, 5] = True
        params = O.init_params(np.random.default_rng(seed + 1), [cfg.F] * cfg.P, cfg.C)
        run_tile_case(tile, cfg, params, mode)
    base.barrier()
    torch.cuda.synchronize()
    if base.rank == 0:
        print("DIST_CHECK_OK world=%d partition=tile" % W, flush=True)
    sys.stdout.flush()
    os._exit(0)


def main():
    if os.environ.get("HAN_DIST_PARTITION", "row") == "tile":
        return main_tile()
    shard = hd.RowShard.init_process_group()
    run_dropout_case(shard)
    run_stacked_case(shard)
    for seed, n, mode in ((61, 257, "reference"), (62, 400, "paper"), (63, 96, "reference")):
        cfg = synth.tiny(seed=seed, n=n, f=36, p=3, deg=6.0)
        cfg.masks[1][:, 5] = True                  # one source every node attends to (crosses shards)
        params = O.init_params(np.random.default_rng(seed + 1), [cfg.F] * cfg.P, cfg.C)
        run_case(shard, cfg, params, mode)
    shard.barrier()
    torch.cuda.synchronize()
    if shard.rank == 0:
        mode = shard.comm if (shard.use_multicast and shard._tables) else "nccl"
        print("DIST_CHECK_OK world=%d comm=%s" % (shard.world, mode), flush=True)
    sys.stdout.flush()
    os._exit(0)     # skip communicator / symmetric-memory teardown (can block at interpreter exit)


if __name__ == "__main__":
    main()
