"""Drop-in counterpart of the reference's ``models/base_gattn.py`` pieces the HAN driver uses:
masked softmax cross-entropy (:41-48), masked accuracy (:61-69) and ``training`` (:12-24: L2 on
every trainable variable + Adam with TF1 semantics).  These are the stock-PyTorch boundary of the
fwd+bwd step; the custom kernels sit below ``inference``.
"""
from __future__ import annotations

import math
from typing import Iterable, List

import torch


class AdamTF1:
    """tf.train.AdamOptimizer (models/base_gattn.py:19): lr_t = lr*sqrt(1-b2^t)/(1-b1^t),
    p -= lr_t * m / (sqrt(v) + eps)  (epsilon outside the bias correction, unlike torch.optim.Adam)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float, beta1=0.9, beta2=0.999, eps=1e-8):
        self.params: List[torch.nn.Parameter] = [p for p in params]
        self.lr, self.beta1, self.beta2, self.eps = lr, beta1, beta2, eps
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]
        self.t = 0

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    @torch.no_grad()
    def step(self):
        self.t += 1
        lr_t = self.lr * math.sqrt(1.0 - self.beta2 ** self.t) / (1.0 - self.beta1 ** self.t)
        ps = [p for p in self.params if p.grad is not None]
        gs = [p.grad for p in ps]
        ms = [self.m[i] for i, p in enumerate(self.params) if p.grad is not None]
        vs = [self.v[i] for i, p in enumerate(self.params) if p.grad is not None]
        torch._foreach_mul_(ms, self.beta1)
        torch._foreach_add_(ms, gs, alpha=1 - self.beta1)
        torch._foreach_mul_(vs, self.beta2)
        torch._foreach_addcmul_(vs, gs, gs, value=1 - self.beta2)
        den = torch._foreach_sqrt(vs)
        torch._foreach_add_(den, self.eps)
        torch._foreach_addcdiv_(ps, ms, den, value=-lr_t)


class TrainOp:
    """What ``training(loss, lr, l2_coef)`` returns in place of a TF train_op."""

    def __init__(self, params, lr, l2_coef):
        self.params = [p for p in params]
        self.l2_coef = l2_coef
        self.opt = AdamTF1(self.params, lr)

    def l2_loss(self) -> torch.Tensor:
        # models/base_gattn.py:14-16: tf.nn.l2_loss(v) = sum(v^2)/2 over ALL trainable variables
        # (the name filter there never matches a real variable name)
        sq = torch._foreach_norm(self.params)
        return torch.stack(sq).pow(2).sum() * (0.5 * self.l2_coef)

    def run(self, loss: torch.Tensor) -> torch.Tensor:
        """One ``sess.run(train_op)``: backward of loss + L2, then the Adam update."""
        self.opt.zero_grad()
        total = loss + self.l2_loss()
        total.backward()
        self.opt.step()
        return total


class BaseGAttN:
    @staticmethod
    def masked_softmax_cross_entropy(logits, labels, mask):
        """models/base_gattn.py:41-48.  logits (N,C); labels one-hot (N,C); mask (N,)."""
        labels = labels.to(logits.dtype)
        loss = -(labels * torch.log_softmax(logits, dim=-1)).sum(-1)     # :43-44
        mask = mask.to(logits.dtype)                                      # :45
        mask = mask / mask.mean()                                         # :46
        return (loss * mask).mean()                                       # :47-48

    @staticmethod
    def masked_accuracy(logits, labels, mask):
        """models/base_gattn.py:61-69."""
        correct = (logits.argmax(1) == labels.argmax(1)).to(logits.dtype)
        mask = mask.to(logits.dtype)
        mask = mask / mask.mean()
        return (correct * mask).mean()

    @staticmethod
    def training(params, lr, l2_coef) -> TrainOp:
        """models/base_gattn.py:12-24.  TF's version takes the loss tensor of a static graph; in eager
        mode the variables are passed instead and the loss is handed to ``TrainOp.run`` each step."""
        if isinstance(params, torch.nn.Module):
            params = params.parameters()
        return TrainOp(params, lr, l2_coef)
