"""CPU oracle for the HAN attention hot path.  TEST INFRASTRUCTURE ONLY.

PINNED TO THE REFERENCE'S OWN SOURCE.  The reference (CG-Labs/HAN, TF1) ships no tests, golden vectors or
seeds, and TensorFlow 1.x cannot be installed in this image; what CAN run here is the reference's code itself:
``utils/process.py`` as shipped (pure numpy) and ``utils/layers.py`` / ``models/gat.py`` /
``models/base_gattn.py`` UNMODIFIED through a small TF1 API shim (``oracle/refrun/tf1_shim.py``: each tf.* op
they call, with its documented semantics, on torch-CPU tensors).  ``oracle/refrun/make_ref_golden.py`` executes
them on seeded inputs and commits the results as ``tests/golden/ref_*.npz``; ``tests/test_oracle_ref.py``
asserts every function below reproduces those fixtures to <= 1e-12 in fp64 (adj_to_bias bit for bit):
outputs, loss, every gradient, the variables after one ``training()`` step, attention coefficients, the three
dropout sites given the same keep masks, TF1 variable names in creation order.  What stays un-pinned is the
arithmetic INSIDE each TF primitive (Eigen's summation order vs torch's) -- bounded by the fp32-vs-fp64
fixture pair at <= 2e-6 -- and ``mode="paper"`` of the semantic layer, which the reference does not implement
(han.pdf Eq. 7-9 only).  Further pins: closed-form known answers, fp64 ``gradcheck``, dense path == edge-list
twin to ~1e-15 (tests/test_oracle.py).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` leg may import this module.  Nothing under ``han_b200/``
does: the product path is CUDA-only and fails loudly without its extension.

Everything here is plain dense torch-CPU in the dtype of its inputs (fp64 =
gold, fp32 = "what TF would compute up to summation order").  Gradients come
from torch autograd, mirroring TF's graph autodiff
(``models/base_gattn.py:22`` ``opt.minimize``), so they are independent of the
hand-derived backward kernels they are used to check.

All file:line citations are relative to /root/reference.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

LEAKY_SLOPE = 0.2  # tf.nn.leaky_relu default alpha (utils/layers.py:27)


# ----------------------------------------------------------------------------
# utils/process.py:14-25
# ----------------------------------------------------------------------------
def adj_to_bias(adj: np.ndarray, sizes: Sequence[int], nhood: int = 1) -> np.ndarray:
    """Restates ``adj_to_bias`` (utils/process.py:14-25), float64 like numpy's default.

    :16 ``mt = np.empty(adj.shape)``; :18 identity; :19-20 ``mt @ (adj + I)``
    repeated ``nhood`` times; :21-24 entries ``> 0`` inside ``sizes[g]**2`` become
    1.0, every other entry keeps its value; :25 ``-1e9 * (1 - mt)``.
    The i,j Python double loop is vectorised; it has no cross-iteration dependency.
    """
    adj = np.asarray(adj)
    nb_graphs = adj.shape[0]
    n = adj.shape[1]
    mt = np.empty(adj.shape)  # float64
    eye = np.eye(n)
    for g in range(nb_graphs):
        mt[g] = eye
        for _ in range(nhood):
            mt[g] = np.matmul(mt[g], (adj[g] + eye))
        s = sizes[g]
        blk = mt[g][:s, :s]
        blk[blk > 0.0] = 1.0
    return -1e9 * (1.0 - mt)


def adj_to_bias_loop(adj: np.ndarray, sizes: Sequence[int], nhood: int = 1) -> np.ndarray:
    """The literal double loop of utils/process.py:21-24 (small inputs only)."""
    adj = np.asarray(adj)
    nb_graphs = adj.shape[0]
    mt = np.empty(adj.shape)
    for g in range(nb_graphs):
        mt[g] = np.eye(adj.shape[1])
        for _ in range(nhood):
            mt[g] = np.matmul(mt[g], (adj[g] + np.eye(adj.shape[1])))
        for i in range(sizes[g]):
            for j in range(sizes[g]):
                if mt[g][i][j] > 0.0:
                    mt[g][i][j] = 1.0
    return -1e9 * (1.0 - mt)


def bias_to_csr(bias: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """The information content of a reference bias matrix: where it is 0.

    After the fp32 cast at the placeholder (ex_acm3025.py:127) an entry takes part
    in the softmax iff its bias is (+/-)0; ``-1e9`` (or lower) is absorbed to an
    exact 0 coefficient.  Returns (indptr int64, indices int32)
    in ``np.nonzero`` (row-major, ascending column) order.
    """
    b = np.asarray(bias)
    if b.ndim == 3:
        assert b.shape[0] == 1
        b = b[0]
    rows, cols = np.nonzero(b == 0)
    n = b.shape[0]
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(indptr, rows + 1, 1)
    indptr = np.cumsum(indptr)
    return indptr.astype(np.int64), cols.astype(np.int32)


# ----------------------------------------------------------------------------
# helpers that restate the TF ops the reference calls
# ----------------------------------------------------------------------------
def _conv1d_k1(x: torch.Tensor, kernel: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """``tf.layers.conv1d(x, out, 1)`` with kernel size 1 is a per-position matmul."""
    y = torch.matmul(x, kernel)
    if bias is not None:
        y = y + bias
    return y


def _dropout(x: torch.Tensor, keep_prob: float, gen: Optional[torch.Generator]) -> torch.Tensor:
    """tf.nn.dropout(x, keep): x / keep * floor(keep + U[0,1))."""
    if keep_prob >= 1.0:
        return x
    u = torch.rand(x.shape, generator=gen, dtype=x.dtype)
    return x / keep_prob * torch.floor(keep_prob + u)


def elu(x: torch.Tensor) -> torch.Tensor:
    return F.elu(x)


def identity(x: torch.Tensor) -> torch.Tensor:
    return x


# ----------------------------------------------------------------------------
# utils/layers.py:7-46
# ----------------------------------------------------------------------------
def attn_head(seq: torch.Tensor, out_sz: int, bias_mat: torch.Tensor, activation: Callable,
              hp: Dict[str, torch.Tensor], in_drop: float = 0.0, coef_drop: float = 0.0,
              residual: bool = False, return_coef: bool = False,
              gen: Optional[torch.Generator] = None, masks: Optional[Dict[str, torch.Tensor]] = None):
    """Restates ``attn_head`` (utils/layers.py:7-46).

    ``seq`` (1,N,F); ``bias_mat`` (1,N,N); ``hp`` holds the variables the reference
    creates implicitly in this call: ``W`` (F,H) [:20 conv1d kernel (1,F,H), no bias],
    ``a1`` (H,), ``b1`` () [:23], ``a2`` (H,), ``b2`` () [:24], ``bias`` (H,) [:35].
    """
    assert hp["W"].shape[1] == out_sz

    def drop(x, rate, which):
        # ``masks`` is a test hook: the three tf.nn.dropout calls take GIVEN 0/1 keep masks (x: (N,F),
        # coef: (N,N), s: (N,H)) instead of fresh random ones, so a CUDA run (or the reference itself, run
        # through the TF1 shim with the same masks) can be compared exactly
        if masks is not None:
            return x * masks[which].to(x.dtype) / (1.0 - rate)
        return _dropout(x, 1.0 - rate, gen)
    if in_drop != 0.0:                                                     # :18-19
        seq = drop(seq, in_drop, "x")
    seq_fts = _conv1d_k1(seq, hp["W"], None)                               # :20  (1,N,H)
    f_1 = _conv1d_k1(seq_fts, hp["a1"].reshape(-1, 1), hp["b1"])           # :23  (1,N,1)
    f_2 = _conv1d_k1(seq_fts, hp["a2"].reshape(-1, 1), hp["b2"])           # :24  (1,N,1)
    logits = f_1 + f_2.transpose(1, 2)                                     # :26  (1,N,N)
    coefs = torch.softmax(F.leaky_relu(logits, LEAKY_SLOPE) + bias_mat, dim=-1)  # :27
    if coef_drop != 0.0:                                                   # :29-30
        coefs = drop(coefs, coef_drop, "coef")
    if in_drop != 0.0:                                                     # :31-32
        seq_fts = drop(seq_fts, in_drop, "s")
    vals = torch.matmul(coefs, seq_fts)                                    # :34
    ret = vals + hp["bias"]                                                # :35
    if residual:                                                           # :38-42
        if seq.shape[-1] != ret.shape[-1]:
            ret = ret + _conv1d_k1(seq, hp["W_res"], hp.get("b_res"))      # :40  (reads the DROPPED seq of :19)
        else:
            seq_fts = ret + seq                                            # :42 (dead store)
    if return_coef:                                                        # :43-44
        return activation(ret), coefs
    return activation(ret)                                                 # :46


def sp_attn_head(seq: torch.Tensor, out_sz: int, adj_rows: np.ndarray, adj_cols: np.ndarray, adj_vals: torch.Tensor,
                 activation: Callable, nb_nodes: int, hp: Dict[str, torch.Tensor], residual: bool = False) -> torch.Tensor:
    """Restates ``sp_attn_head`` (utils/layers.py:85-127), dropout off, on the stored entries
    (row, col, value) of the batch-1 ``tf.SparseTensor`` adjacency.

    :95-96 ``logits = sparse_add(adj * f_1, adj * f_2^T)``: stored value w_ij scales BOTH score terms,
    ``l_ij = w_ij f1_i + w_ij f2_j``; :97-99 leaky_relu on the stored values; :100 ``tf.sparse_softmax`` =
    softmax over each row's STORED entries (an explicit zero still takes part); :113-115 sparse @ dense.
    """
    assert hp["W"].shape[1] == out_sz
    seq_fts = _conv1d_k1(seq, hp["W"], None)                               # :90   (1,N,H)
    f_1 = _conv1d_k1(seq_fts, hp["a1"].reshape(-1, 1), hp["b1"])[0, :, 0]  # :93
    f_2 = _conv1d_k1(seq_fts, hp["a2"].reshape(-1, 1), hp["b2"])[0, :, 0]  # :94
    rows = torch.from_numpy(np.asarray(adj_rows, dtype=np.int64))
    cols = torch.from_numpy(np.asarray(adj_cols, dtype=np.int64))
    w = adj_vals.to(seq.dtype)
    e = F.leaky_relu(w * f_1[rows] + w * f_2[cols], LEAKY_SLOPE)           # :95-99
    m = torch.full((nb_nodes,), -math.inf, dtype=seq.dtype).scatter_reduce(0, rows, e.detach(), reduce="amax")
    ex = torch.exp(e - m[rows])                                            # :100
    den = torch.zeros(nb_nodes, dtype=seq.dtype).index_add(0, rows, ex)
    coefs = ex / den[rows]
    S = seq_fts[0]                                                         # :112 squeeze
    vals = torch.zeros_like(S).index_add(0, rows, coefs.unsqueeze(1) * S[cols])   # :113
    ret = vals.unsqueeze(0) + hp["bias"]                                   # :114-116
    if residual and seq.shape[-1] != ret.shape[-1]:                        # :119-121
        ret = ret + _conv1d_k1(seq, hp["W_res"], hp.get("b_res"))
    return activation(ret)                                                 # :127


def attn_head_const_1(seq: torch.Tensor, out_sz: int, bias_mat: torch.Tensor, activation: Callable,
                      hp: Dict[str, torch.Tensor], residual: bool = False) -> torch.Tensor:
    """Restates ``attn_head_const_1`` (utils/layers.py:49-81), dropout off: the logits are the 0/1
    adjacency recovered from the bias (:55), so every neighbour weighs 1/deg.  ``hp``: W (F,H), bias (H,)."""
    adj_mat = 1.0 - bias_mat / -1e9                                          # :55
    seq_fts = _conv1d_k1(seq, hp["W"], None)                                 # :59
    coefs = torch.softmax(F.leaky_relu(adj_mat, LEAKY_SLOPE) + bias_mat, dim=-1)   # :62-63
    ret = torch.matmul(coefs, seq_fts) + hp["bias"]                          # :70-71
    if residual and seq.shape[-1] != ret.shape[-1]:                          # :74-76
        ret = ret + _conv1d_k1(seq, hp["W_res"], hp.get("b_res"))
    return activation(ret)                                                   # :81


# ----------------------------------------------------------------------------
# utils/layers.py:132-164
# ----------------------------------------------------------------------------
def SimpleAttLayer(inputs: torch.Tensor, attention_size: int, sp: Dict[str, torch.Tensor],
                   time_major: bool = False, return_alphas: bool = False, mode: str = "reference"):
    """Restates ``SimpleAttLayer`` (utils/layers.py:132-164).

    ``inputs`` (N,P,D); ``sp``: ``w_omega`` (D,A), ``b_omega`` (A,), ``u_omega`` (A,)
    [:145-147].  ``mode="reference"`` is the shipped per-node softmax over
    meta-paths (:156); ``mode="paper"`` is han.pdf Eq. 7-9 (node-mean first).
    """
    if isinstance(inputs, tuple):                                          # :134-136
        inputs = torch.cat(inputs, 2)
    if time_major:                                                         # :138-140
        inputs = inputs.transpose(0, 1)
    assert sp["w_omega"].shape == (inputs.shape[2], attention_size)
    v = torch.tanh(torch.tensordot(inputs, sp["w_omega"], dims=1) + sp["b_omega"])  # :152
    vu = torch.tensordot(v, sp["u_omega"], dims=1)                         # :155 (N,P)
    if mode == "reference":
        alphas = torch.softmax(vu, dim=-1)                                 # :156
    elif mode == "paper":
        beta = torch.softmax(vu.mean(dim=0), dim=-1)                       # han.pdf Eq. 8
        alphas = beta.unsqueeze(0).expand_as(vu)
    else:
        raise ValueError(mode)
    output = torch.sum(inputs * alphas.unsqueeze(-1), 1)                   # :159
    if not return_alphas:
        return output
    return output, alphas


# ----------------------------------------------------------------------------
# models/gat.py:34-77
# ----------------------------------------------------------------------------
def head_params(params: Dict, p: int, k: int, layer: int = 0) -> Dict[str, torch.Tensor]:
    """Slice head ``k`` of meta-path ``p`` out of the pre-concatenated layout.

    Layout (shared with the product so tests can hand the same tensors to both):
    ``params['W'][p]`` (F,K*H) with head k in columns k*H:(k+1)*H, ``a1/a2[p]`` (K,H),
    ``b1/b2[p]`` (K,), ``bias[p]`` (K*H,).
    """
    src = params if layer == 0 else params["deep"][layer - 1]      # stacked layers keep the same keys
    H = src["a1"][p].shape[1]
    sl = slice(k * H, (k + 1) * H)
    hp = {"W": src["W"][p][:, sl], "a1": src["a1"][p][k], "b1": src["b1"][p][k],
          "a2": src["a2"][p][k], "b2": src["b2"][p][k], "bias": src["bias"][p][sl]}
    if "W_res" in src:
        hp["W_res"], hp["b_res"] = src["W_res"][p][:, sl], src["b_res"][p][sl]
    return hp


def HeteGAT_multi_inference(inputs_list, nb_classes, nb_nodes, training, attn_drop, ffd_drop,
                            bias_mat_list, hid_units, n_heads, params: Dict,
                            activation: Callable = elu, residual: bool = False,
                            mp_att_size: int = 128, semantic_mode: str = "reference",
                            return_coef: bool = False, gen: Optional[torch.Generator] = None):
    """Restates ``HeteGAT_multi.inference`` (models/gat.py:35-77).

    Returns ``(logits (1,N,C), final_embed (N,D), att_val (N,P))`` (:76-77); with
    ``return_coef`` additionally the per-(meta-path, head) dense coefficient list.
    """
    embed_list = []
    coef_list = []
    for p, (inputs, bias_mat) in enumerate(zip(inputs_list, bias_mat_list)):     # :39
        attns = []
        for k in range(n_heads[0]):                                              # :42
            r = attn_head(inputs, hid_units[0], bias_mat, activation, head_params(params, p, k),
                          in_drop=ffd_drop, coef_drop=attn_drop, residual=False,
                          return_coef=return_coef, gen=gen)                      # :43-45
            if return_coef:
                attns.append(r[0]); coef_list.append(r[1])
            else:
                attns.append(r)
        h_1 = torch.cat(attns, dim=-1)                                           # :46  (1,N,D)
        for i in range(1, len(hid_units)):                                       # :48-57
            attns = []
            for k in range(n_heads[i]):
                r = attn_head(h_1, hid_units[i], bias_mat, activation, head_params(params, p, k, layer=i),
                              in_drop=ffd_drop, coef_drop=attn_drop, residual=residual,
                              return_coef=return_coef, gen=gen)                  # :52-56
                if return_coef:
                    attns.append(r[0]); coef_list.append(r[1])
                else:
                    attns.append(r)
            h_1 = torch.cat(attns, dim=-1)                                       # :57
        embed_list.append(h_1.squeeze(0).unsqueeze(1))                           # :58  (N,1,D)
    multi_embed = torch.cat(embed_list, dim=1)                                   # :60  (N,P,D)
    final_embed, att_val = SimpleAttLayer(multi_embed, mp_att_size, params, time_major=False,
                                          return_alphas=True, mode=semantic_mode)  # :61-63
    out = []
    for i in range(n_heads[-1]):                                                 # :66-68
        out.append(torch.matmul(final_embed, params["Wc"][i]) + params["bc"][i])
    logits = sum(out) / n_heads[-1]                                              # :72
    logits = logits.unsqueeze(0)                                                 # :76
    if return_coef:
        return logits, final_embed, att_val, coef_list
    return logits, final_embed, att_val


def GAT_inference(inputs, nb_classes, nb_nodes, training, attn_drop, ffd_drop, bias_mat, hid_units, n_heads,
                  params: Dict, activation: Callable = elu, residual: bool = False,
                  gen: Optional[torch.Generator] = None):
    """Restates ``GAT.inference`` (models/gat.py:9-32).  ``params``: {"hidden": [layer dicts], "out": layer dict},
    each layer dict in the concatenated-heads layout of ``head_params`` (W (F,K*H), a1/a2 (K,H), b1/b2 (K,), bias)."""
    def head(lay, k, H):
        sl = slice(k * H, (k + 1) * H)
        hp = {"W": lay["W"][:, sl], "a1": lay["a1"][k], "b1": lay["b1"][k], "a2": lay["a2"][k], "b2": lay["b2"][k],
              "bias": lay["bias"][sl]}
        if "W_res" in lay:
            hp["W_res"], hp["b_res"] = lay["W_res"][:, sl], lay["b_res"][sl]
        return hp
    h_1 = inputs
    for i in range(len(hid_units)):                                              # :11-23
        attns = [attn_head(h_1, hid_units[i], bias_mat, activation, head(params["hidden"][i], k, hid_units[i]),
                           in_drop=ffd_drop, coef_drop=attn_drop, residual=residual and i >= 1, gen=gen)
                 for k in range(n_heads[i])]
        h_1 = torch.cat(attns, dim=-1)
    out = [attn_head(h_1, nb_classes, bias_mat, identity, head(params["out"], k, nb_classes),
                     in_drop=ffd_drop, coef_drop=attn_drop, residual=False, gen=gen)
           for k in range(n_heads[-1])]                                          # :25-29
    return sum(out) / n_heads[-1]                                                # :30


def init_gat_params(rng: np.random.Generator, ft_size: int, nb_classes: int, hid_units: Sequence[int],
                    n_heads: Sequence[int], dtype=torch.float64, residual: bool = False) -> Dict:
    def glorot(fan_in, fan_out, shape):
        lim = math.sqrt(6.0 / (fan_in + fan_out))
        return torch.from_numpy(rng.uniform(-lim, lim, size=shape)).to(dtype)

    def small(shape):
        return torch.from_numpy(rng.normal(0, 0.1, size=shape)).to(dtype)

    def layer(F, K, H, res):
        lay = {"W": torch.cat([glorot(F, H, (F, H)) for _ in range(K)], dim=1), "a1": glorot(H, 1, (K, H)),
               "b1": small((K,)), "a2": glorot(H, 1, (K, H)), "b2": small((K,)), "bias": small((K * H,))}
        if res:
            lay["W_res"] = torch.cat([glorot(F, H, (F, H)) for _ in range(K)], dim=1)
            lay["b_res"] = small((K * H,))
        return lay
    hidden, F = [], ft_size
    for l, H in enumerate(hid_units):
        hidden.append(layer(F, n_heads[l], H, residual and l >= 1 and F != H))
        F = n_heads[l] * H
    return {"hidden": hidden, "out": layer(F, n_heads[-1], nb_classes, False)}


# ----------------------------------------------------------------------------
# models/base_gattn.py
# ----------------------------------------------------------------------------
def masked_softmax_cross_entropy(logits: torch.Tensor, labels: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """models/base_gattn.py:41-48.  ``labels`` one-hot (N,C) (cast to float, no grad), ``mask`` (N,)."""
    labels = labels.to(logits.dtype)
    loss = -(labels * torch.log_softmax(logits, dim=-1)).sum(-1)                 # :43-44
    mask = mask.to(logits.dtype)                                                 # :45
    mask = mask / mask.mean()                                                    # :46
    loss = loss * mask                                                           # :47
    return loss.mean()                                                           # :48


def masked_accuracy(logits: torch.Tensor, labels: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """models/base_gattn.py:61-69."""
    correct = (logits.argmax(1) == labels.argmax(1)).to(logits.dtype)            # :63-65
    mask = mask.to(logits.dtype)
    mask = mask / mask.mean()                                                    # :67
    return (correct * mask).mean()                                               # :68-69


def flat_params(params: Dict) -> List[torch.Tensor]:
    out = []
    for key in ("W", "a1", "b1", "a2", "b2", "bias", "Wc", "bc"):
        out.extend(params[key])
    out.extend([params["w_omega"], params["b_omega"], params["u_omega"]])
    for lay in params.get("deep", []):
        for v in lay.values():
            out.extend(v)
    return out


def l2_loss_all(params: Dict, l2_coef: float) -> torch.Tensor:
    """models/base_gattn.py:14-16: tf.nn.l2_loss = sum(v**2)/2 over ALL trainable variables
    (the name filter at :15-16 never matches a real TF variable name)."""
    return sum((v * v).sum() / 2 for v in flat_params(params)) * l2_coef


def adam_step_tf1(param: torch.Tensor, grad: torch.Tensor, m: torch.Tensor, v: torch.Tensor, t: int,
                  lr: float = 0.005, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8):
    """tf.train.AdamOptimizer update (models/base_gattn.py:19-22) [external: TF1 semantics]:
    lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m,v EMAs; p -= lr_t * m / (sqrt(v) + eps)."""
    lr_t = lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
    m = beta1 * m + (1 - beta1) * grad
    v = beta2 * v + (1 - beta2) * grad * grad
    param = param - lr_t * m / (v.sqrt() + eps)
    return param, m, v


def step_loss(inputs_list, bias_mat_list, labels, mask, params, nb_classes, hid_units, n_heads,
              l2_coef: float = 0.001, semantic_mode: str = "reference", residual: bool = False):
    """One forward of the training objective the driver builds (ex_acm3025.py:139-152):
    inference -> reshape -> masked CE -> + L2.  Returns (total, ce, logits, final_embed, att_val)."""
    logits, final_embed, att_val = HeteGAT_multi_inference(
        inputs_list, nb_classes, inputs_list[0].shape[1], True, 0.0, 0.0, bias_mat_list,
        hid_units, n_heads, params, mp_att_size=params["w_omega"].shape[1], semantic_mode=semantic_mode,
        residual=residual)
    log_resh = logits.reshape(-1, nb_classes)                                    # ex_acm3025.py:146
    ce = masked_softmax_cross_entropy(log_resh, labels.reshape(-1, nb_classes), mask.reshape(-1))
    total = ce + l2_loss_all(params, l2_coef)
    return total, ce, logits, final_embed, att_val


# ----------------------------------------------------------------------------
# Edge-list twin (for graphs whose N^2 does not fit): same math restricted to edges.
# ----------------------------------------------------------------------------
def attn_head_edges(x: torch.Tensor, indptr: np.ndarray, indices: np.ndarray, hp: Dict[str, torch.Tensor],
                    activation: Callable = elu, return_coef: bool = False):
    """Edge-list restatement of utils/layers.py:20-46 with dropout off: softmax over the
    neighbours N(i) = CSR row i only.  ``x`` (N,F) -> (N,H) [+ alpha (E,)]."""
    n = x.shape[0]
    S = x @ hp["W"]
    f1 = S @ hp["a1"] + hp["b1"]
    f2 = S @ hp["a2"] + hp["b2"]
    deg = np.diff(indptr)
    rows = torch.from_numpy(np.repeat(np.arange(n, dtype=np.int64), deg))
    cols = torch.from_numpy(np.asarray(indices, dtype=np.int64))
    e = F.leaky_relu(f1[rows] + f2[cols], LEAKY_SLOPE)
    m = torch.full((n,), -math.inf, dtype=x.dtype).scatter_reduce(0, rows, e, reduce="amax")
    ex = torch.exp(e - m[rows])
    den = torch.zeros(n, dtype=x.dtype).index_add(0, rows, ex)
    alpha = ex / den[rows]
    vals = torch.zeros_like(S).index_add(0, rows, alpha.unsqueeze(1) * S[cols])
    out = activation(vals + hp["bias"])
    if return_coef:
        return out, alpha
    return out


def inference_edges(x_list, csr_list, params, n_heads, hid_units, mp_att_size=128,
                    activation: Callable = elu, semantic_mode: str = "reference"):
    """Edge-list twin of ``HeteGAT_multi_inference`` (dropout off)."""
    embeds = []
    for p, (x, (indptr, indices)) in enumerate(zip(x_list, csr_list)):
        heads = [attn_head_edges(x, indptr, indices, head_params(params, p, k), activation)
                 for k in range(n_heads[0])]
        embeds.append(torch.cat(heads, dim=-1).unsqueeze(1))
    Z = torch.cat(embeds, dim=1)
    final_embed, att_val = SimpleAttLayer(Z, mp_att_size, params, return_alphas=True, mode=semantic_mode)
    out = [final_embed @ params["Wc"][i] + params["bc"][i] for i in range(n_heads[-1])]
    logits = (sum(out) / n_heads[-1]).unsqueeze(0)
    return logits, final_embed, att_val


# ----------------------------------------------------------------------------
# Parameter initialisation from a seeded numpy generator.
# ----------------------------------------------------------------------------
def init_params(rng: np.random.Generator, ft_sizes: Sequence[int], nb_classes: int, hid: int = 8,
                heads: int = 8, mp_att_size: int = 128, out_heads: int = 1,
                dtype=torch.float64, zero_bias: bool = False, deep: Sequence[Tuple[int, int]] = (),
                residual: bool = False) -> Dict:
    """Glorot-uniform conv1d/dense kernels, N(0,0.1^2) semantic variables.  The reference
    zero-initialises b1,b2,bias,bc; ``zero_bias=False`` draws them small-random instead so
    parity tests exercise them."""
    D = hid * heads

    def glorot(fan_in, fan_out, shape):
        lim = math.sqrt(6.0 / (fan_in + fan_out))
        return torch.from_numpy(rng.uniform(-lim, lim, size=shape)).to(dtype)

    def small(shape):
        if zero_bias:
            return torch.zeros(shape, dtype=dtype)
        return torch.from_numpy(rng.normal(0, 0.1, size=shape)).to(dtype)

    params = {"W": [], "a1": [], "b1": [], "a2": [], "b2": [], "bias": []}
    for Fp in ft_sizes:
        params["W"].append(torch.cat([glorot(Fp, hid, (Fp, hid)) for _ in range(heads)], dim=1))
        params["a1"].append(glorot(hid, 1, (heads, hid)))
        params["b1"].append(small((heads,)))
        params["a2"].append(glorot(hid, 1, (heads, hid)))
        params["b2"].append(small((heads,)))
        params["bias"].append(small((D,)))
    Fl = D
    for (Kl, Hl) in deep:            # stacked layers (heads, hid) after the first, models/gat.py:48-57
        lay = {k: [] for k in ("W", "a1", "b1", "a2", "b2", "bias")}
        if residual and Fl != Hl:
            lay["W_res"], lay["b_res"] = [], []
        for _ in ft_sizes:
            lay["W"].append(torch.cat([glorot(Fl, Hl, (Fl, Hl)) for _ in range(Kl)], dim=1))
            lay["a1"].append(glorot(Hl, 1, (Kl, Hl)))
            lay["b1"].append(small((Kl,)))
            lay["a2"].append(glorot(Hl, 1, (Kl, Hl)))
            lay["b2"].append(small((Kl,)))
            lay["bias"].append(small((Kl * Hl,)))
            if "W_res" in lay:
                lay["W_res"].append(torch.cat([glorot(Fl, Hl, (Fl, Hl)) for _ in range(Kl)], dim=1))
                lay["b_res"].append(small((Kl * Hl,)))
        params.setdefault("deep", []).append(lay)
        Fl = Kl * Hl
    D = Fl
    params["w_omega"] = torch.from_numpy(rng.normal(0, 0.1, size=(D, mp_att_size))).to(dtype)
    params["b_omega"] = torch.from_numpy(rng.normal(0, 0.1, size=(mp_att_size,))).to(dtype)
    params["u_omega"] = torch.from_numpy(rng.normal(0, 0.1, size=(mp_att_size,))).to(dtype)
    params["Wc"] = [glorot(D, nb_classes, (D, nb_classes)) for _ in range(out_heads)]
    params["bc"] = [small((nb_classes,)) for _ in range(out_heads)]
    return params


def params_to(params: Dict, dtype=None, requires_grad: bool = False) -> Dict:
    def conv(t):
        t = t.detach().clone()
        if dtype is not None:
            t = t.to(dtype)
        return t.requires_grad_(requires_grad)
    out = {}
    for k, v in params.items():
        if k == "deep":
            out[k] = [{kk: [conv(t) for t in vv] for kk, vv in lay.items()} for lay in v]
        else:
            out[k] = [conv(t) for t in v] if isinstance(v, list) else conv(v)
    return out
