"""CPU timing legs built on the oracle.  TEST / BENCH INFRASTRUCTURE ONLY (see han_oracle.py header):
imported only by bench.py's ``cpu_baseline`` leg and ``--impl reference`` arm.

The reference itself (TF1) cannot run in this image, so "the reference's CPU path" is the fp32 dense
restatement in han_oracle.py, run on torch-CPU with every host thread (kind = "port").

Two samples of a workload:
  * small configs (ACM/DBLP/IMDB-shaped): the complete reference step -- adj_to_bias output fed as
    dense fp32 (1,N,N) biases, P*K attn_head calls, SimpleAttLayer, dense, masked CE + L2, autograd
    backward -- exactly what ex_acm3025.py:139-152,190 runs per step.
  * large configs (2M-node, MAG-scale): one N x N tensor is terabytes, so the sample is a block of R
    destination rows against ALL N source columns: the dense chain of utils/layers.py:26-35,46
    (f1 + f2^T, leaky_relu, + bias_mat, softmax, coefs @ seq_fts, bias_add, elu) forward + autograd
    backward for all P*K heads.  Projection / semantic / classifier (O(N*F*D), <1% of the chain at
    this N) are left out, which favours the CPU.  edges/s = edges in those R rows / time.
"""
from __future__ import annotations

import os
import time
from typing import List

import numpy as np
import torch
import torch.nn.functional as F

from . import han_oracle as O


def host_threads() -> int:
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        pass
    return n


def dense_full_step_seconds(cfg, params32, steps: int, warmup: int, hid_units=(8,), n_heads=(8, 1)) -> List[float]:
    """Times `steps` complete dense fp32 reference steps (fwd + bwd) on a SmallConfig."""
    torch.set_num_threads(host_threads())
    X = torch.from_numpy(cfg.X).unsqueeze(0)
    biases = [torch.from_numpy(O.adj_to_bias(a, [cfg.N], 1)).float() for a in cfg.adjs()]
    labels = torch.from_numpy(cfg.labels)
    mask = torch.from_numpy(cfg.train_mask.astype(np.float32))
    times = []
    for it in range(warmup + steps):
        p = O.params_to(params32, torch.float32, requires_grad=True)
        t0 = time.perf_counter()
        total, *_ = O.step_loss([X] * cfg.P, biases, labels, mask, p, cfg.C, list(hid_units), list(n_heads))
        total.backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return times


def make_rowblock_problem(N: int, P: int, K: int, H: int, R: int, mean_degree: float, seed: int):
    """Synthetic stand-ins of the right shape for the row-block sample: per meta-path S (N,K*H),
    f2 (N,K), for R rows f1 (R,K) and the CSR rows (self-loop + random sources)."""
    g = torch.Generator().manual_seed(seed)
    prob = []
    d = int(round(mean_degree))
    for p in range(P):
        S = torch.randn(N, K * H, generator=g)
        f2 = torch.randn(N, K, generator=g)
        f1 = torch.randn(R, K, generator=g)
        cols = torch.randint(0, N, (R, d), generator=g)
        cols[:, 0] = torch.arange(R)            # self-loops (rows 0..R-1 are the sample's nodes)
        # the bias rows are an INPUT of the step (utils/process.py:25 runs once per graph, ex_acm3025.py:118)
        bias_mat = torch.full((R, N), -1e9)
        bias_mat.scatter_(1, cols, 0.0)
        prob.append((S, f2, f1, bias_mat))
    return prob


def rowblock_edges(prob) -> int:
    return int(sum(int((b == 0).sum()) for _, _, _, b in prob))


def dense_rowblock_step_seconds(prob, K: int, H: int) -> float:
    """One fwd+bwd of the reference's dense chain for R rows x N columns, all P*K heads (inputs prebuilt)."""
    torch.set_num_threads(host_threads())
    t0 = time.perf_counter()
    for S, f2, f1, bias_mat in prob:
        for k in range(K):
            seq_fts = S[:, k * H:(k + 1) * H].clone().requires_grad_(True)
            f_1 = f1[:, k:k + 1].clone().requires_grad_(True)
            f_2 = f2[:, k:k + 1].clone().requires_grad_(True)
            logits = f_1 + f_2.transpose(0, 1)                            # layers.py:26
            coefs = torch.softmax(F.leaky_relu(logits, 0.2) + bias_mat, dim=-1)   # :27
            vals = coefs @ seq_fts                                        # :34
            ret = F.elu(vals)                                             # :35,46 (zero bias)
            ret.sum().backward()
    return time.perf_counter() - t0
