"""the ablation / alternative operators and the homogeneous GAT, against
the dense fp64 oracle -- attn_head_const_1 (utils/layers.py:49-81), sp_attn_head (:85-127), the residual
branch of attn_head (:38-42) and GAT.inference (models/gat.py:8-32)."""
import numpy as np
import pytest
import torch

from han_b200 import synth
from oracle import han_oracle as O
from tests.util import assert_close

pytestmark = pytest.mark.gpu


def _head_inputs(seed, n=130, f=22, h=8):
    cfg = synth.tiny(seed=seed, n=n, f=f, p=1, deg=6.0)
    rng = np.random.default_rng(seed + 1)
    lim = np.sqrt(6.0 / (f + h))
    hp = {"W": torch.from_numpy(rng.uniform(-lim, lim, (f, h))), "a1": torch.from_numpy(rng.normal(size=h)),
          "b1": torch.tensor(0.05, dtype=torch.float64), "a2": torch.from_numpy(rng.normal(size=h)),
          "b2": torch.tensor(-0.03, dtype=torch.float64), "bias": torch.from_numpy(rng.normal(0, 0.1, h)),
          "W_res": torch.from_numpy(rng.uniform(-lim, lim, (f, h))), "b_res": torch.from_numpy(rng.normal(0, 0.1, h))}
    bias = torch.from_numpy(O.adj_to_bias(cfg.adjs()[0], [cfg.N], 1))
    X = torch.from_numpy(cfg.X).double()[None]
    return cfg, hp, bias, X


def _cuda_params(hp, keys):
    return {k: torch.nn.Parameter(hp[k].float().cuda()) for k in keys}


def _check_grads(pp, po, keys):
    for k in keys:
        assert_close(pp[k].grad, po[k].grad, "d" + k)


@pytest.mark.parametrize("residual", [False, True])
def test_attn_head_const_1_is_the_neighbour_mean(residual):
    import han_b200 as hb
    cfg, hp, bias, X = _head_inputs(101)
    keys = ["W", "bias"] + (["W_res", "b_res"] if residual else [])
    po = {k: hp[k].clone().requires_grad_(True) for k in keys}
    out_o = O.attn_head_const_1(X, 8, bias, O.elu, po, residual=residual)
    g = torch.from_numpy(np.random.default_rng(5).normal(size=tuple(out_o.shape)))
    (out_o * g).sum().backward()
    pp = _cuda_params(hp, keys)
    graph = hb.process.adj_to_bias(cfg.adjs()[0], [cfg.N])
    out_p = hb.layers.attn_head_const_1(X.float().cuda(), 8, graph, hb.layers.elu, residual=residual, params=pp)
    (out_p * g.float().cuda()).sum().backward()
    assert_close(out_p, out_o, "out")
    _check_grads(pp, po, keys)
    if not residual:       # uniform weights: pre-activation = mean of the neighbours' projected features + bias
        m = torch.from_numpy(cfg.masks[0]).double()
        mean = (m / m.sum(1, keepdim=True)) @ (X[0] @ hp["W"]) + hp["bias"]
        assert_close(out_p, O.elu(mean)[None], "neighbour mean")


def test_attn_head_residual_branch_matches_oracle():
    import han_b200 as hb
    cfg, hp, bias, X = _head_inputs(111)
    keys = ["W", "a1", "b1", "a2", "b2", "bias", "W_res", "b_res"]
    po = {k: hp[k].clone().requires_grad_(True) for k in keys}
    out_o = O.attn_head(X, 8, bias, O.elu, po, residual=True)
    g = torch.from_numpy(np.random.default_rng(6).normal(size=tuple(out_o.shape)))
    (out_o * g).sum().backward()
    pp = _cuda_params(hp, keys)
    graph = hb.process.adj_to_bias(cfg.adjs()[0], [cfg.N])
    out_p = hb.layers.attn_head(X.float().cuda(), 8, graph, hb.layers.elu, residual=True, params=pp)
    (out_p * g.float().cuda()).sum().backward()
    assert_close(out_p, out_o, "out")
    _check_grads(pp, po, keys)
    # equal widths: the reference's residual is a dead store (layers.py:42) -> identical to residual=False
    cfg2, hp2, bias2, X2 = _head_inputs(112, f=8, h=8)
    p2 = _cuda_params(hp2, ["W", "a1", "b1", "a2", "b2", "bias"])
    g2 = hb.process.adj_to_bias(cfg2.adjs()[0], [cfg2.N])
    with torch.no_grad():
        a = hb.layers.attn_head(X2.float().cuda(), 8, g2, hb.layers.elu, residual=True, params=p2)
        b = hb.layers.attn_head(X2.float().cuda(), 8, g2, hb.layers.elu, residual=False, params=p2)
    assert torch.equal(a, b) and "W_res" not in p2


def test_sp_attn_head_on_a_binary_sparse_adjacency_equals_attn_head():
    import han_b200 as hb
    cfg, hp, bias, X = _head_inputs(121)
    keys = ["W", "a1", "b1", "a2", "b2", "bias"]
    pp = _cuda_params(hp, keys)
    m = torch.from_numpy(cfg.masks[0])
    sp = m.float().to_sparse_coo().cuda()
    Xc = X.float().cuda()
    with torch.no_grad():
        a = hb.layers.sp_attn_head(Xc, 8, sp, hb.layers.elu, cfg.N, params=pp)
        b = hb.layers.attn_head(Xc, 8, hb.process.adj_to_bias(cfg.adjs()[0], [cfg.N]), hb.layers.elu, params=pp)
        o = O.attn_head(X, 8, bias, O.elu, hp)
    assert torch.equal(a, b)
    assert_close(a, o, "sp_attn_head vs dense oracle")
    with torch.no_grad():      # stored values != 1 scale the logits (utils/layers.py:95-96): a different result
        w2 = hb.layers.sp_attn_head(Xc, 8, (m.float() * 2.0).to_sparse_coo().cuda(), hb.layers.elu, cfg.N, params=pp)
        rows, cols = np.nonzero(cfg.masks[0])
        ow = O.sp_attn_head(X, 8, rows, cols, torch.full((len(rows),), 2.0, dtype=torch.float64), O.elu, cfg.N, hp)
    assert not torch.equal(w2, a)
    assert_close(w2, ow, "sp_attn_head with weights 2 vs oracle")


@pytest.mark.parametrize("hid_units,n_heads,residual,classes", [((8,), (4, 1), False, 3), ((8, 8), (2, 2, 2), True, 7)])
def test_gat_inference_matches_oracle(hid_units, n_heads, residual, classes):
    """models/gat.py:8-32 incl. the averaged output heads of width nb_classes (padded to the kernels' H)."""
    import han_b200 as hb
    cfg = synth.tiny(seed=131, n=140, f=24, p=1, c=classes, deg=6.0)
    params = O.init_gat_params(np.random.default_rng(132), cfg.F, cfg.C, hid_units, n_heads, residual=residual)
    po = {"hidden": [{k: v.clone().requires_grad_(True) for k, v in lay.items()} for lay in params["hidden"]],
          "out": {k: v.clone().requires_grad_(True) for k, v in params["out"].items()}}
    X = torch.from_numpy(cfg.X).double()[None]
    bias = torch.from_numpy(O.adj_to_bias(cfg.adjs()[0], [cfg.N], 1))
    lo = O.GAT_inference(X, cfg.C, cfg.N, False, 0.0, 0.0, bias, list(hid_units), list(n_heads), po, residual=residual)
    labels = torch.from_numpy(cfg.labels).double()
    mask = torch.from_numpy(cfg.train_mask.astype(np.float64))
    O.masked_softmax_cross_entropy(lo.reshape(-1, cfg.C), labels, mask).backward()
    gp = hb.GATParams(cfg.F, cfg.C, hid_units, n_heads, device="cuda", residual=residual).load_dict(params)
    graph = hb.process.adj_to_bias(cfg.adjs()[0], [cfg.N])
    lp = hb.GAT.inference(X.float().cuda(), cfg.C, cfg.N, False, 0.0, 0.0, graph, list(hid_units), list(n_heads),
                          residual=residual, params=gp)
    assert lp.shape == (1, cfg.N, cfg.C)
    hb.BaseGAttN.masked_softmax_cross_entropy(lp.reshape(-1, cfg.C), labels.float().cuda(), mask.float().cuda()).backward()
    assert_close(lp, lo, "logits")
    grads = gp.grad_dict()
    for l, lay in enumerate(po["hidden"]):
        for k, v in lay.items():
            assert_close(grads["hidden"][l][k], v.grad, f"hidden[{l}].d{k}")
    for k, v in po["out"].items():
        assert_close(grads["out"][k], v.grad, f"out.d{k}")
    # training mode runs (dropout 0.6 on inputs and coefficients of every layer)
    if not residual:
        lt = hb.GAT.inference(X.float().cuda(), cfg.C, cfg.N, True, 0.6, 0.6, graph, list(hid_units), list(n_heads), params=gp)
        lt.sum().backward()
        assert torch.isfinite(lt).all() and not torch.equal(lt, lp)
