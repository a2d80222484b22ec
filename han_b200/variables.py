"""Parameter store standing in for TF1's implicit variable creation.

The reference creates fresh ``tf.Variable``s inside every ``attn_head`` / ``SimpleAttLayer`` /
``tf.layers.dense`` call (utils/layers.py:20,23,24,35,145-147; models/gat.py:68), in call order,
inside the default graph.  ``HANParams`` holds the same tensors (same shapes, same initialisers,
SURVEY.md Appendix B) pre-concatenated per meta-path, and ``tf_variable_names`` gives the TF1
auto-names so a reference checkpoint could be mapped onto it.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
from torch import nn


def _glorot_(t: torch.Tensor, fan_in: int, fan_out: int, gen=None) -> torch.Tensor:
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return t.uniform_(-lim, lim, generator=gen)


class HANParams(nn.Module):
    """All trainable variables of ``HeteGAT_multi.inference`` for P meta-paths.

    W[p] (F_p, K*H): head k in columns k*H:(k+1)*H == conv1d kernel (1,F,H) of head k (layers.py:20)
    a1[p], a2[p] (K,H); b1[p], b2[p] (K,): the two 1-channel conv1d's (layers.py:23-24)
    bias[p] (K*H,): contrib.layers.bias_add (layers.py:35)
    w_omega (D,A), b_omega (A,), u_omega (A,): SimpleAttLayer (layers.py:145-147)
    Wc[i] (D,C), bc[i] (C,): tf.layers.dense heads (gat.py:66-68)

    Stacked layers (``len(hid_units) > 1``, gat.py:48-57): layer l >= 1 of meta-path p reads the previous
    layer's concatenated heads (F_l = K_{l-1} H_{l-1}); its variables live in ``deep[l-1]`` under the same
    keys (lists over p), plus ``W_res[p] (F_l, K_l H_l)`` / ``b_res[p] (K_l H_l,)`` when ``residual`` -- the
    per-head ``conv1d(seq, H_l, 1)`` of layers.py:40, which has a bias.  D above is the LAST layer's width.
    """

    def __init__(self, ft_sizes: Sequence[int], nb_classes: int, hid_units: Sequence[int] = (8,),
                 n_heads: Sequence[int] = (8, 1), mp_att_size: int = 128, device=None,
                 generator: Optional[torch.Generator] = None, residual: bool = False):
        super().__init__()
        if len(n_heads) < len(hid_units) + 1:
            raise ValueError("n_heads needs one entry per attention layer plus the output-layer entry")
        self.P = len(ft_sizes)
        self.K, self.H = int(n_heads[0]), int(hid_units[0])
        self.layer_dims = [(int(n_heads[l]), int(hid_units[l])) for l in range(len(hid_units))]
        self.D = self.layer_dims[-1][0] * self.layer_dims[-1][1]
        self.A, self.C = int(mp_att_size), int(nb_classes)
        self.out_heads = int(n_heads[-1])
        self.residual = bool(residual)
        K, H, D = self.K, self.H, self.K * self.H
        kw = dict(dtype=torch.float32, device="cpu")  # init on host for device-independent streams
        g = generator

        def P_(t):
            return nn.Parameter(t.to(device) if device is not None else t)

        self.W = nn.ParameterList()
        self.a1, self.b1, self.a2, self.b2, self.bias = (nn.ParameterList() for _ in range(5))
        for F in ft_sizes:
            Wp = torch.empty(F, D, **kw)
            for k in range(K):  # one glorot draw per head: fan_in=F, fan_out=H
                _glorot_(Wp[:, k * H:(k + 1) * H], F, H, g)
            self.W.append(P_(Wp))
            self.a1.append(P_(_glorot_(torch.empty(K, H, **kw), H, 1, g)))
            self.b1.append(P_(torch.zeros(K, **kw)))
            self.a2.append(P_(_glorot_(torch.empty(K, H, **kw), H, 1, g)))
            self.b2.append(P_(torch.zeros(K, **kw)))
            self.bias.append(P_(torch.zeros(D, **kw)))
        self.deep = nn.ModuleList()
        for l in range(1, len(self.layer_dims)):
            Kl, Hl = self.layer_dims[l]
            Fl = self.layer_dims[l - 1][0] * self.layer_dims[l - 1][1]
            lay = nn.ModuleDict({k: nn.ParameterList() for k in ("W", "a1", "b1", "a2", "b2", "bias")})
            if self.residual and Fl != Hl:                                   # layers.py:39-40
                lay["W_res"], lay["b_res"] = nn.ParameterList(), nn.ParameterList()
            for _ in ft_sizes:
                Wp = torch.empty(Fl, Kl * Hl, **kw)
                for k in range(Kl):
                    _glorot_(Wp[:, k * Hl:(k + 1) * Hl], Fl, Hl, g)
                lay["W"].append(P_(Wp))
                lay["a1"].append(P_(_glorot_(torch.empty(Kl, Hl, **kw), Hl, 1, g)))
                lay["b1"].append(P_(torch.zeros(Kl, **kw)))
                lay["a2"].append(P_(_glorot_(torch.empty(Kl, Hl, **kw), Hl, 1, g)))
                lay["b2"].append(P_(torch.zeros(Kl, **kw)))
                lay["bias"].append(P_(torch.zeros(Kl * Hl, **kw)))
                if "W_res" in lay:
                    Wr = torch.empty(Fl, Kl * Hl, **kw)
                    for k in range(Kl):
                        _glorot_(Wr[:, k * Hl:(k + 1) * Hl], Fl, Hl, g)
                    lay["W_res"].append(P_(Wr))
                    lay["b_res"].append(P_(torch.zeros(Kl * Hl, **kw)))
            self.deep.append(lay)
        D = self.D
        self.w_omega = P_(torch.empty(D, self.A, **kw).normal_(0, 0.1, generator=g))
        self.b_omega = P_(torch.empty(self.A, **kw).normal_(0, 0.1, generator=g))
        self.u_omega = P_(torch.empty(self.A, **kw).normal_(0, 0.1, generator=g))
        self.Wc = nn.ParameterList([P_(_glorot_(torch.empty(D, self.C, **kw), D, self.C, g))
                                    for _ in range(self.out_heads)])
        self.bc = nn.ParameterList([P_(torch.zeros(self.C, **kw)) for _ in range(self.out_heads)])
        # dropout seed word (device-resident int32; bumped once per training forward)
        s0 = int(torch.randint(0, 2 ** 31 - 1, (1,), generator=g).item())
        seed = torch.tensor([s0], dtype=torch.int32)
        self.register_buffer("drop_seed", seed.to(device) if device is not None else seed)

    # ---- exchange with the oracle's dict layout (tests) -----------------------------------------
    def load_dict(self, params: Dict) -> "HANParams":
        with torch.no_grad():
            for key in ("W", "a1", "b1", "a2", "b2", "bias", "Wc", "bc"):
                for dst, src in zip(getattr(self, key), params[key]):
                    dst.copy_(src.to(dst.dtype))
            for key in ("w_omega", "b_omega", "u_omega"):
                getattr(self, key).copy_(params[key].to(torch.float32))
            for lay, src in zip(self.deep, params.get("deep", [])):
                for key, plist in lay.items():
                    for dst, t in zip(plist, src[key]):
                        dst.copy_(t.to(dst.dtype))
        return self

    def to_dict(self) -> Dict:
        d = {key: [t.detach() for t in getattr(self, key)]
             for key in ("W", "a1", "b1", "a2", "b2", "bias", "Wc", "bc")}
        for key in ("w_omega", "b_omega", "u_omega"):
            d[key] = getattr(self, key).detach()
        if len(self.deep):
            d["deep"] = [{key: [t.detach() for t in plist] for key, plist in lay.items()} for lay in self.deep]
        return d

    def grad_dict(self) -> Dict:
        d = {key: [t.grad for t in getattr(self, key)]
             for key in ("W", "a1", "b1", "a2", "b2", "bias", "Wc", "bc")}
        for key in ("w_omega", "b_omega", "u_omega"):
            d[key] = getattr(self, key).grad
        if len(self.deep):
            d["deep"] = [{key: [t.grad for t in plist] for key, plist in lay.items()} for lay in self.deep]
        return d

    def tf_variable_names(self) -> Dict[str, str]:
        """TF1 auto-names in creation order (meta-path major, head minor), Appendix B."""
        names = {}

        def sfx(i):
            return "" if i == 0 else f"_{i}"

        for p in range(self.P):
            for k in range(self.K):
                c = 3 * (p * self.K + k)
                i = p * self.K + k
                names[f"W[{p}][:, head {k}]"] = f"conv1d{sfx(c)}/kernel"
                names[f"a1[{p}][{k}]"] = f"conv1d{sfx(c + 1)}/kernel"
                names[f"b1[{p}][{k}]"] = f"conv1d{sfx(c + 1)}/bias"
                names[f"a2[{p}][{k}]"] = f"conv1d{sfx(c + 2)}/kernel"
                names[f"b2[{p}][{k}]"] = f"conv1d{sfx(c + 2)}/bias"
                names[f"bias[{p}][head {k}]"] = f"BiasAdd{sfx(i)}/biases"
        names["w_omega"], names["b_omega"], names["u_omega"] = "Variable", "Variable_1", "Variable_2"
        for i in range(self.out_heads):
            names[f"Wc[{i}]"] = f"dense{sfx(i)}/kernel"
            names[f"bc[{i}]"] = f"dense{sfx(i)}/bias"
        return names


def _head_layer(F: int, K: int, H: int, gen, device) -> nn.ParameterDict:
    """The variables K ``attn_head(seq (.,F), out_sz=H)`` calls create (layers.py:20,23,24,35), concatenated."""
    W = torch.empty(F, K * H)
    for k in range(K):
        _glorot_(W[:, k * H:(k + 1) * H], F, H, gen)
    mk = lambda t: nn.Parameter(t.to(device) if device is not None else t)
    return nn.ParameterDict({"W": mk(W), "a1": mk(_glorot_(torch.empty(K, H), H, 1, gen)), "b1": mk(torch.zeros(K)),
                             "a2": mk(_glorot_(torch.empty(K, H), H, 1, gen)), "b2": mk(torch.zeros(K)),
                             "bias": mk(torch.zeros(K * H))})


class GATParams(nn.Module):
    """Variables of the homogeneous ``GAT.inference`` (models/gat.py:8-32): ``hidden[l]`` for the
    ``len(hid_units)`` attention layers (K_l = n_heads[l] heads of hid_units[l]), ``out`` for the
    ``n_heads[-1]`` output heads of width nb_classes (:25-29).  Keys as in ``HANParams``."""

    def __init__(self, ft_size: int, nb_classes: int, hid_units: Sequence[int] = (8,), n_heads: Sequence[int] = (8, 1),
                 device=None, generator: Optional[torch.Generator] = None, residual: bool = False):
        super().__init__()
        self.C, self.out_heads = int(nb_classes), int(n_heads[-1])
        self.layer_dims = [(int(n_heads[l]), int(hid_units[l])) for l in range(len(hid_units))]
        self.hidden = nn.ModuleList()
        F = int(ft_size)
        for l, (K, H) in enumerate(self.layer_dims):
            lay = _head_layer(F, K, H, generator, device)
            if residual and l >= 1 and F != H:                              # gat.py:20 passes residual from layer 1 on
                r = _head_layer(F, K, H, generator, device)
                lay["W_res"], lay["b_res"] = r["W"], r["bias"]
            self.hidden.append(lay)
            F = K * H
        self.out = _head_layer(F, self.out_heads, self.C, generator, device)
        s0 = int(torch.randint(0, 2 ** 31 - 1, (1,), generator=generator).item())
        seed = torch.tensor([s0], dtype=torch.int32)
        self.register_buffer("drop_seed", seed.to(device) if device is not None else seed)

    def load_dict(self, params: Dict) -> "GATParams":
        with torch.no_grad():
            for lay, src in zip(list(self.hidden) + [self.out], list(params["hidden"]) + [params["out"]]):
                for key, dst in lay.items():
                    dst.copy_(src[key].to(dst.dtype))
        return self

    def grad_dict(self) -> Dict:
        return {"hidden": [{k: v.grad for k, v in lay.items()} for lay in self.hidden],
                "out": {k: v.grad for k, v in self.out.items()}}


# The reference builds its variables in TF's process-wide default graph; mirror that with a
# process-wide default store that `inference` creates on first use and reuses afterwards.
_default_store: Optional[HANParams] = None


def get_default_store() -> Optional[HANParams]:
    return _default_store


def set_default_store(p: Optional[HANParams]) -> None:
    global _default_store
    _default_store = p


def reset_default_graph() -> None:
    """tf.reset_default_graph() analogue: forget the implicitly created variables."""
    set_default_store(None)
