"""Training-loop protocol of the reference driver (ex_acm3025.py:161-270) on top of the CUDA hot path.

One "epoch" is one full-graph step (``batch_size = 1`` graph, :21,171): a training forward/backward/update
with both dropouts at 0.6 (:185-186), then a validation forward with dropout off (:207-209).  Early
stopping follows :225-239 exactly -- the patience counter resets when validation accuracy OR loss
improves, the checkpoint is written only when BOTH do -- and the test pass runs on the restored
checkpoint (:247-270).  With ``cuda_graph=True`` the training step (forward + backward + fused L2/Adam
update, dropout seeds advancing on the device) and the evaluation forward are each captured once and
replayed, so an epoch costs two graph launches and one small device->host read.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence

import torch

from . import layers, variables
from .base_gattn import BaseGAttN
from .gat import HeteGAT_multi
from .graphs import GraphedStep


@dataclass
class FitResult:
    epochs_run: int
    stopped_early: bool
    best_val_loss: float
    best_val_acc: float
    checkpoint_val_loss: float
    checkpoint_val_acc: float
    test_loss: float
    test_acc: float
    final_embed: torch.Tensor            # (N, D) of the restored checkpoint, dropout off (:262)
    att_val: torch.Tensor                # (N, P) semantic attention of the test pass
    history: List[dict] = field(default_factory=list)


class EarlyStopping:
    """The stopping rule of ex_acm3025.py:225-239 as a small state machine.  ``update`` returns
    ``(save, stop)``: the patience counter resets when validation accuracy OR loss is at least as good as
    the best seen, a checkpoint is due only when BOTH are, and ``stop`` turns true after ``patience``
    consecutive epochs with neither."""

    def __init__(self, patience: int):
        self.patience = patience
        self.vlss_mn, self.vacc_mx, self.curr_step = float("inf"), 0.0, 0
        self.ck_loss = self.ck_acc = float("nan")

    def update(self, vl_loss: float, vl_acc: float):
        save = False
        if vl_acc >= self.vacc_mx or vl_loss <= self.vlss_mn:               # :225
            if vl_acc >= self.vacc_mx and vl_loss <= self.vlss_mn:          # :226-229
                self.ck_acc, self.ck_loss, save = vl_acc, vl_loss, True
            self.vacc_mx, self.vlss_mn = max(vl_acc, self.vacc_mx), min(vl_loss, self.vlss_mn)   # :230-231
            self.curr_step = 0
            return save, False
        self.curr_step += 1
        return False, self.curr_step == self.patience                      # :233-239


def _as_device(x, dev, dtype=torch.float32):
    t = x if isinstance(x, torch.Tensor) else torch.as_tensor(x)
    return t.to(device=dev, dtype=dtype)


def fit(fea_list: Sequence, biases_list: Sequence, y_train, y_val, y_test, train_mask, val_mask, test_mask, *,
        nb_epochs: int = 200, patience: int = 100, lr: float = 0.005, l2_coef: float = 0.001,
        hid_units: Sequence[int] = (8,), n_heads: Sequence[int] = (8, 1), mp_att_size: int = 128,
        attn_drop: float = 0.6, ffd_drop: float = 0.6, checkpt_file: Optional[str] = None,
        params: Optional[variables.HANParams] = None, model=HeteGAT_multi, project_mode: int = 0,
        cuda_graph: bool = True, log: Optional[Callable[[str], None]] = print, log_every: int = 1,
        device=None) -> FitResult:
    """fea_list: P arrays (1,N,F) or (N,F); biases_list: P ``MetaPathGraph``s (``process.adj_to_bias``) or
    dense reference biases; y_*: one-hot (1,N,C) or (N,C) with the rows outside the split zeroed
    (ex_acm3025.py:72-77); *_mask: (1,N) or (N,)."""
    dev = torch.device("cuda" if device is None else device)
    xs = [_as_device(x, dev) for x in fea_list]
    xs = [x if x.dim() == 3 else x.unsqueeze(0) for x in xs]
    if all(torch.equal(xs[0], x) for x in xs[1:]) and len(xs) > 1:
        xs = [xs[0]] * len(xs)                      # ACM feeds one matrix P times (:86): one projection launch
    graphs = [layers.as_graph(b, dev) for b in biases_list]
    for g in graphs:
        g.transpose()
    N, C = xs[0].shape[1], int(torch.as_tensor(y_train).shape[-1])
    ys = [_as_device(y, dev).reshape(N, C) for y in (y_train, y_val, y_test)]
    ms = [_as_device(m, dev).reshape(N) for m in (train_mask, val_mask, test_mask)]
    hid_units, n_heads = list(hid_units), list(n_heads)
    if params is None:
        params = variables.HANParams([x.shape[2] for x in xs], C, hid_units, n_heads, mp_att_size, device=dev)
    train_op = BaseGAttN.training(params, lr, l2_coef)                     # :152

    def forward(split: int, drop_a: float, drop_f: float):
        logits, emb, att = model.inference(xs, C, N, drop_a > 0, drop_a, drop_f, graphs, hid_units, n_heads,
                                           mp_att_size=mp_att_size, params=params, project_mode=project_mode)
        lg = logits.reshape(-1, C)                                          # :146-148
        return (BaseGAttN.masked_softmax_cross_entropy(lg, ys[split], ms[split]),
                BaseGAttN.masked_accuracy(lg, ys[split], ms[split]), emb, att)

    def train_step():
        loss, acc, _, att = forward(0, attn_drop, ffd_drop)
        train_op.run(loss)
        return torch.stack([loss.detach(), acc.detach()]), att.detach().mean(0)

    def eval_step(split: int):
        with torch.no_grad():
            loss, acc, emb, att = forward(split, 0.0, 0.0)
        return torch.stack([loss, acc]), emb, att

    val_step = lambda: eval_step(1)
    if cuda_graph:
        train_step = _graphed_after_reset(train_step, params, train_op)
        val_step = GraphedStep(val_step, warmup=1)

    rule = EarlyStopping(patience)
    best_state = None
    history, stopped, epoch = [], False, -1
    for epoch in range(nb_epochs):
        tr, att_mean = train_step()
        vl, _, _ = val_step()
        both = torch.cat([tr, vl, att_mean]).tolist()                      # the epoch's one device->host read
        tr_loss, tr_acc, vl_loss, vl_acc = both[:4]
        history.append({"epoch": epoch, "train_loss": tr_loss, "train_acc": tr_acc, "val_loss": vl_loss,
                        "val_acc": vl_acc, "att_val": both[4:]})
        if log is not None and epoch % log_every == 0:
            log(f"Epoch: {epoch}, att_val: {both[4:]}")
            log("Training: loss = %.5f, acc = %.5f | Val: loss = %.5f, acc = %.5f" % (tr_loss, tr_acc, vl_loss, vl_acc))
        save, stopped = rule.update(vl_loss, vl_acc)
        if save:
            best_state = {k: v.detach().clone() for k, v in params.state_dict().items()}
            if checkpt_file is not None:
                torch.save(best_state, checkpt_file)
        if stopped:
            if log is not None:
                log(f"Early stop! Min loss: {rule.vlss_mn}, Max accuracy: {rule.vacc_mx}")
                log(f"Early stop model validation loss: {rule.ck_loss}, accuracy: {rule.ck_acc}")
            break

    if best_state is not None:                                              # :247
        with torch.no_grad():
            for k, v in params.state_dict().items():
                v.copy_(best_state[k])                                      # in place: the flat buffer keeps its address
    ts, emb, att = eval_step(2)
    ts_loss, ts_acc = ts.tolist()
    if log is not None:
        log(f"Test loss: {ts_loss}; Test accuracy: {ts_acc}")
    return FitResult(epoch + 1, stopped, rule.vlss_mn, rule.vacc_mx, rule.ck_loss, rule.ck_acc, ts_loss, ts_acc, emb.detach(), att.detach(),
                     history)


def _graphed_after_reset(train_step, params, train_op):
    """Capture the training step without letting the warm-up/capture runs train the model: GraphedStep
    runs the step eagerly before capturing, so the variables, Adam moments, step counter and dropout
    seed are put back afterwards (all in place -- the captured graph holds their addresses)."""
    train_op.opt.flatten()
    opt = train_op.opt
    saved = [t.clone() for t in (opt.flat_p, opt.m, opt.v, opt.t, params.drop_seed)]
    g = GraphedStep(train_step, warmup=1)
    with torch.no_grad():
        for dst, src in zip((opt.flat_p, opt.m, opt.v, opt.t, params.drop_seed), saved):
            dst.copy_(src)
    return g
