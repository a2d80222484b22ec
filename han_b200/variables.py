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
    """

    def __init__(self, ft_sizes: Sequence[int], nb_classes: int, hid_units: Sequence[int] = (8,),
                 n_heads: Sequence[int] = (8, 1), mp_att_size: int = 128, device=None,
                 generator: Optional[torch.Generator] = None):
        super().__init__()
        if len(hid_units) != 1:
            raise NotImplementedError("stacked attention layers (models/gat.py:48-57) are not built yet")
        self.P = len(ft_sizes)
        self.K, self.H = int(n_heads[0]), int(hid_units[0])
        self.D = self.K * self.H
        self.A, self.C = int(mp_att_size), int(nb_classes)
        self.out_heads = int(n_heads[-1])
        K, H, D = self.K, self.H, self.D
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
        return self

    def to_dict(self) -> Dict:
        d = {key: [t.detach() for t in getattr(self, key)]
             for key in ("W", "a1", "b1", "a2", "b2", "bias", "Wc", "bc")}
        for key in ("w_omega", "b_omega", "u_omega"):
            d[key] = getattr(self, key).detach()
        return d

    def grad_dict(self) -> Dict:
        d = {key: [t.grad for t in getattr(self, key)]
             for key in ("W", "a1", "b1", "a2", "b2", "bias", "Wc", "bc")}
        for key in ("w_omega", "b_omega", "u_omega"):
            d[key] = getattr(self, key).grad
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
