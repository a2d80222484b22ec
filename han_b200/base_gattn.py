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


class FusedAdamTF1:
    """The same update as ``AdamTF1`` plus the L2 term of ``training`` (models/base_gattn.py:14-16), as
    ONE kernel launch (``han_adam_l2_step``) over a flat buffer that holds every variable.

    ``flatten()`` (called on first use) moves the parameters into the flat buffer -- each ``p.data``
    becomes a view at a 256-byte aligned offset -- and gives every parameter a ``.grad`` view of a flat
    gradient buffer, which autograd then accumulates into in place.  Sharded runs all-reduce that one
    buffer.  The step counter is a device word, so forward + backward + update capture into one CUDA graph.
    """
    ALIGN = 64   # floats: keeps every view 256-byte aligned (TMA descriptors need 16)

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float, l2_coef: float,
                 beta1=0.9, beta2=0.999, eps=1e-8):
        self.params: List[torch.nn.Parameter] = [p for p in params]
        self.lr, self.l2_coef, self.beta1, self.beta2, self.eps = lr, l2_coef, beta1, beta2, eps
        self.flat_p = None

    def flatten(self):
        if self.flat_p is not None:
            return
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedAdamTF1 needs CUDA parameters (han_adam_l2_step has no host twin)")
        offs, n = [], 0
        for p in self.params:
            offs.append(n)
            n += -(-p.numel() // self.ALIGN) * self.ALIGN
        self.flat_p = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros_like(self.flat_p)
        self.m = torch.zeros_like(self.flat_p)
        self.v = torch.zeros_like(self.flat_p)
        self.t = torch.zeros(1, dtype=torch.int32, device=dev)
        self._gviews = []
        with torch.no_grad():
            for p, o in zip(self.params, offs):
                view = self.flat_p[o:o + p.numel()].view(p.shape)
                view.copy_(p.detach())
                p.data = view
                self._gviews.append(self.flat_g[o:o + p.numel()].view(p.shape))

    def zero_grad(self):
        self.flatten()
        self.flat_g.zero_()
        for p, g in zip(self.params, self._gviews):
            p.grad = g

    @torch.no_grad()
    def step(self):
        from . import _lib
        self.t.add_(1)
        _lib.call("han_adam_l2_step", _lib.ptr(self.flat_p), _lib.ptr(self.flat_g), _lib.ptr(self.m),
                  _lib.ptr(self.v), self.flat_p.numel(), _lib.ptr(self.t), self.lr, self.beta1, self.beta2,
                  self.eps, self.l2_coef, _lib.stream_ptr())


class TrainOp:
    """What ``training(loss, lr, l2_coef)`` returns in place of a TF train_op.

    ``run(loss)`` is one ``sess.run(train_op)``.  On CUDA parameters the L2 gradient and the Adam update are
    one fused launch (``FusedAdamTF1``) and ``run`` returns the loss it was given (the reference fetches the
    cross-entropy, not the L2-augmented objective, ex_acm3025.py:190); ``l2_loss()`` stays available as an
    autograd term for callers that want the objective itself (the parity tests and ``bench.py``)."""

    def __init__(self, params, lr, l2_coef, fused=None):
        self.params = [p for p in params]
        self.l2_coef = l2_coef
        self.fused = all(p.is_cuda for p in self.params) if fused is None else fused
        self.opt = FusedAdamTF1(self.params, lr, l2_coef) if self.fused else AdamTF1(self.params, lr)

    def l2_loss(self) -> torch.Tensor:
        # models/base_gattn.py:14-16: tf.nn.l2_loss(v) = sum(v^2)/2 over ALL trainable variables
        # (the name filter there never matches a real variable name)
        sq = torch._foreach_norm(self.params)
        return torch.stack(sq).pow(2).sum() * (0.5 * self.l2_coef)

    def run(self, loss: torch.Tensor, dist=None) -> torch.Tensor:
        """One ``sess.run(train_op)``: backward of loss (+ L2), then the Adam update.  ``dist``: the
        row-shard context of a multi-GPU run (gradients are summed over ranks before the update; pass
        this rank's share of the loss, ``dist.masked_loss(..., train_op=None)``)."""
        self.opt.zero_grad()
        if self.fused:
            loss.backward()
            if dist is not None:
                dist.all_reduce_flat(self.opt.flat_g)
            self.opt.step()
            return loss.detach()
        total = loss + self.l2_loss()
        total.backward()
        if dist is not None:
            raise NotImplementedError("sharded training uses the fused optimizer (CUDA parameters)")
        self.opt.step()
        return total.detach()


class BaseGAttN:
    @staticmethod
    def masked_softmax_cross_entropy(logits, labels, mask):
        """models/base_gattn.py:41-48.  logits (N,C); labels one-hot (N,C); mask (N,)."""
        if logits.is_cuda and logits.dtype == torch.float32:
            # one kernel: loss partials + d(loss)/d(logits); mean(loss * mask / mean(mask)) == sum(loss * mask) / sum(mask)
            from . import ops
            mask = mask.to(torch.float32)
            return ops.masked_ce(logits, labels, mask, mask.sum())
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

    # ---- the remaining helpers of the reference class (metric / loss glue outside the hot path; stock ops) ----
    @staticmethod
    def loss(logits, labels, nb_classes, class_weights):
        """models/base_gattn.py:5-10: class-weighted sparse softmax cross-entropy, mean over samples.
        logits (M,C); labels (M,) integer classes; class_weights (C,)."""
        labels = labels.long()
        w = torch.as_tensor(class_weights, dtype=logits.dtype, device=logits.device)[labels]      # one_hot . weights
        xent = torch.nn.functional.cross_entropy(logits, labels, reduction="none")
        return (xent * w).mean()

    @staticmethod
    def preshape(logits, labels, nb_classes):
        """models/base_gattn.py:26-32."""
        return logits.reshape(-1, nb_classes), labels.reshape(-1)

    @staticmethod
    def confmat(logits, labels):
        """models/base_gattn.py:34-36: confusion matrix, rows = true class, columns = predicted class."""
        preds = logits.argmax(1)
        labels = labels.long()
        n = int(max(int(preds.max()), int(labels.max())) + 1) if labels.numel() else 0
        return torch.bincount(labels * n + preds, minlength=n * n).reshape(n, n)

    @staticmethod
    def masked_sigmoid_cross_entropy(logits, labels, mask):
        """models/base_gattn.py:52-62 (multi-label nodes): per-node mean of the element-wise sigmoid
        cross-entropy, then the same mask normalisation as the softmax variant."""
        labels = labels.to(logits.dtype)
        loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, labels, reduction="none").mean(1)
        mask = mask.to(logits.dtype)
        mask = mask / mask.mean()
        return (loss * mask).mean()

    @staticmethod
    def micro_f1(logits, labels, mask):
        """models/base_gattn.py:75-101: micro-averaged F1 of round(sigmoid(logits)) over the masked nodes,
        counted in integers."""
        predicted = torch.round(torch.sigmoid(logits)).to(torch.int64)
        labels = labels.to(torch.int64)
        mask = mask.to(torch.int64).unsqueeze(-1)
        tp = torch.count_nonzero(predicted * labels * mask)
        fp = torch.count_nonzero(predicted * (labels - 1) * mask)
        fn = torch.count_nonzero((predicted - 1) * labels * mask)
        precision = tp / (tp + fp)
        recall = tp / (tp + fn)
        return (2 * precision * recall / (precision + recall)).to(torch.float32)

    @staticmethod
    def training(params, lr, l2_coef) -> TrainOp:
        """models/base_gattn.py:12-24.  TF's version takes the loss tensor of a static graph; in eager
        mode the variables are passed instead and the loss is handed to ``TrainOp.run`` each step."""
        if isinstance(params, torch.nn.Module):
            params = params.parameters()
        return TrainOp(params, lr, l2_coef)
