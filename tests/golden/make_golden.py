"""Generates the golden fixtures under tests/golden/*.npz from the fp64 oracle.

    python tests/golden/make_golden.py

The reference has no golden vectors of its own and cannot run here (no TensorFlow), so these pin the
ORACLE's outputs (inputs, weights, outputs, loss, every gradient) on seeded inputs; the GPU parity
tests compare the CUDA path with them, and tests/test_oracle.py checks the oracle still reproduces
them.  Each case returns a flat dict of numpy arrays.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from han_b200 import synth  # noqa: E402
from oracle import han_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
LIST_KEYS = ("W", "a1", "b1", "a2", "b2", "bias", "Wc", "bc")
VEC_KEYS = ("w_omega", "b_omega", "u_omega")


def flatten_params(params, prefix):
    out = {}
    for k in LIST_KEYS:
        for i, t in enumerate(params[k]):
            out[f"{prefix}{k}{i}"] = t.detach().numpy()
    for k in VEC_KEYS:
        out[f"{prefix}{k}"] = params[k].detach().numpy()
    return out


def unflatten_params(d, prefix, P, out_heads=1):
    params = {k: [torch.from_numpy(d[f"{prefix}{k}{i}"]) for i in range(P)] for k in LIST_KEYS[:6]}
    for k in ("Wc", "bc"):
        params[k] = [torch.from_numpy(d[f"{prefix}{k}{i}"]) for i in range(out_heads)]
    for k in VEC_KEYS:
        params[k] = torch.from_numpy(d[f"{prefix}{k}"])
    return params


def run_case(cfg, heads, hid, att, mode, seed):
    rng = np.random.default_rng(seed)
    params = O.init_params(rng, [cfg.F] * cfg.P, cfg.C, hid=hid, heads=heads, mp_att_size=att)
    p = O.params_to(params, torch.float64, requires_grad=True)
    X = torch.from_numpy(cfg.X).double().unsqueeze(0)
    biases = [torch.from_numpy(O.adj_to_bias(a, [cfg.N], 1)) for a in cfg.adjs()]
    labels = torch.from_numpy(cfg.labels).double()
    mask = torch.from_numpy(cfg.train_mask.astype(np.float64))
    total, ce, logits, fe, av = O.step_loss([X] * cfg.P, biases, labels, mask, p, cfg.C, [hid], [heads, 1],
                                            0.001, mode)
    total.backward()
    grads = {k: ([t.grad for t in v] if isinstance(v, list) else v.grad) for k, v in p.items()}
    out = {"X": cfg.X, "labels": cfg.labels, "train_mask": cfg.train_mask,
           "meta": np.array([cfg.N, cfg.F, cfg.P, cfg.C, heads, hid, att, 1 if mode == "paper" else 0], dtype=np.int64),
           "logits": logits.detach().numpy(), "final_embed": fe.detach().numpy(), "att_val": av.detach().numpy(),
           "ce": ce.detach().numpy(), "total": total.detach().numpy()}
    for i, b in enumerate(biases):
        indptr, indices = O.bias_to_csr(b.numpy())
        out[f"indptr{i}"], out[f"indices{i}"] = indptr, indices
    out.update(flatten_params(params, "p_"))
    out.update(flatten_params(grads, "g_"))
    return out


def case_tiny_p2_k8h8():
    return run_case(synth.tiny(seed=21, n=96, f=40, p=2, c=3, deg=6.0), 8, 8, 128, "reference", 121)


def case_tiny_p3_k4h8_paper():
    return run_case(synth.tiny(seed=22, n=70, f=33, p=3, c=4, deg=4.0), 4, 8, 64, "paper", 122)


def case_degenerate_rows():
    cfg = synth.tiny(seed=23, n=50, f=12, p=2, c=3, deg=5.0, binary=True)
    m0, m1 = cfg.masks
    m0[3, :] = False; m0[3, 3] = True          # single neighbour: alpha = 1
    m0[7, :] = True                            # full row: every node is a neighbour
    m1[:, 11] = True                           # one very popular source (long transposed row)
    m1[20, :] = False; m1[20, 5] = True        # no self-loop, single off-diagonal neighbour
    return run_case(cfg, 8, 8, 128, "reference", 123)


CASES = {"tiny_p2_k8h8": case_tiny_p2_k8h8, "tiny_p3_k4h8_paper": case_tiny_p3_k4h8_paper,
         "degenerate_rows": case_degenerate_rows}


def load_case(name):
    """-> (SmallConfig, params dict fp64, stored arrays) for the GPU parity tests."""
    d = np.load(os.path.join(HERE, name + ".npz"))
    N, F, P, C, heads, hid, att, paper = (int(x) for x in d["meta"])
    masks = []
    for i in range(P):
        indptr, indices = d[f"indptr{i}"], d[f"indices{i}"]
        m = np.zeros((N, N), dtype=bool)
        m[np.repeat(np.arange(N), np.diff(indptr)), indices] = True
        masks.append(m)
    tm = d["train_mask"].astype(bool)
    cfg = synth.SmallConfig(name, N, F, C, [f"MP{i}" for i in range(P)], d["X"], masks, d["labels"], tm, ~tm, ~tm)
    params = unflatten_params(d, "p_", P)
    return cfg, params, d, dict(heads=heads, hid=hid, att=att, mode="paper" if paper else "reference")


if __name__ == "__main__":
    for name, fn in CASES.items():
        arrays = fn()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
        print(name, sum(a.nbytes for a in arrays.values()) // 1024, "KiB")
