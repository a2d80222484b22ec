"""Host-side logic of the training protocol (no GPU): the early-stopping rule of ex_acm3025.py:225-239,
the planted generator, and the embedding-evaluation helpers (jhyexp.py:20-86)."""
import math

import numpy as np

from han_b200 import jhyexps, synth
from han_b200.train import EarlyStopping


def test_early_stopping_resets_on_either_metric_and_saves_only_on_both():
    r = EarlyStopping(patience=2)
    assert r.update(1.0, 0.5) == (True, False)          # first epoch always checkpoints (acc >= 0, loss <= inf)
    assert r.update(0.9, 0.4) == (False, False)         # loss better, acc worse: reset, no checkpoint
    assert (r.vlss_mn, r.vacc_mx, r.curr_step) == (0.9, 0.5, 0)
    assert r.update(0.95, 0.45) == (False, False)       # neither: stale 1
    assert r.update(0.95, 0.5) == (False, False)        # acc ties the best: counts as improvement (>=), loss not -> no save
    assert r.curr_step == 0
    assert r.update(0.9, 0.6) == (True, False)          # both (loss ties with <=)
    assert (r.ck_loss, r.ck_acc) == (0.9, 0.6)
    assert r.update(1.0, 0.1) == (False, False)
    assert r.update(1.0, 0.1) == (False, True)          # patience reached
    assert math.isnan(EarlyStopping(1).ck_loss)


def test_planted_graph_is_homophilous_and_split_is_a_partition():
    cfg = synth.planted(seed=3, n=400, f=90)
    y = cfg.labels.argmax(1)
    pap = cfg.masks[0].copy()
    np.fill_diagonal(pap, False)
    r, c = np.nonzero(pap)
    assert (y[r] == y[c]).mean() > 0.75
    assert (cfg.masks[0] == cfg.masks[0].T).all() and cfg.masks[1].diagonal().all()
    assert (cfg.train_mask.astype(int) + cfg.val_mask + cfg.test_mask == 1).all()
    assert set(np.unique(cfg.X)) == {0.0, 1.0}


def test_knn_and_kmeans_protocol_on_separable_embeddings(capsys):
    rng = np.random.default_rng(0)
    y = rng.integers(0, 3, size=300)
    x = np.eye(3)[y] * 4.0 + rng.normal(size=(300, 3))
    knn = jhyexps.my_KNN(x[None], np.eye(3)[y], time=3, seed=1)
    assert set(knn) == {0.2, 0.4, 0.6, 0.8} and all(ma > 0.9 and mi > 0.9 for ma, mi in knn.values())
    nmi, ari = jhyexps.my_Kmeans(x, y, k=3, time=2, return_NMI=True, seed=1)
    assert nmi > 0.8 and ari > 0.8
    out = capsys.readouterr().out
    assert "KNN(3avg, split:0.2, k=5)" in out and "NMI (2 avg)" in out


def test_split_layout_cuts_heavy_rows_exactly():
    """Index arithmetic of MetaPathGraph.split_view (heavy-row splitting) on CPU tensors:
    empty rows, rows of exactly S / k*S / k*S+1 edges, no heavy row at all."""
    import torch
    from han_b200.graph import split_layout
    rng = np.random.default_rng(0)
    for S, degs in ((4, [0, 1, 4, 5, 8, 9, 0, 3, 17, 4]), (16, list(rng.integers(0, 100, size=200))), (8, [3, 2, 1]),
                    (1, [2, 0, 3])):
        deg = np.asarray(degs, dtype=np.int64)
        indptr = np.concatenate([[0], np.cumsum(deg)])
        iv, vptr, vmap, hrows, hptr, n_slots = split_layout(torch.from_numpy(indptr), S)
        iv, vptr, vmap, hrows, hptr = (t.numpy() for t in (iv, vptr, vmap, hrows, hptr))
        nseg = np.maximum(1, -(-deg // S))
        # every real offset survives, segments are at most S long, lengths per row add up
        assert iv[0] == 0 and iv[-1] == indptr[-1] and (np.diff(iv) <= S).all() and (np.diff(iv) >= 0).all()
        assert np.array_equal(vptr, np.concatenate([[0], np.cumsum(nseg)]))
        assert np.array_equal(iv[vptr[:-1]], indptr[:-1])
        for r in range(len(deg)):
            segs = np.diff(iv[vptr[r]:vptr[r + 1] + 1])
            assert segs.sum() == deg[r] and (vmap[vptr[r]:vptr[r + 1], 0] == r).all()
            if nseg[r] > 1:
                assert (segs[:-1] == S).all() and 1 <= segs[-1] <= S
        # cut rows own consecutive partial slots in row order; whole rows own none
        heavy = np.nonzero(nseg > 1)[0]
        assert np.array_equal(hrows, heavy) and n_slots == nseg[heavy].sum() == hptr[-1]
        assert np.array_equal(hptr, np.concatenate([[0], np.cumsum(nseg[heavy])]))
        slots = vmap[:, 1]
        assert (slots[np.isin(vmap[:, 0], heavy)] == np.arange(n_slots)).all()
        assert (slots[~np.isin(vmap[:, 0], heavy)] == -1).all()


def test_base_gattn_metric_helpers_match_their_definitions():
    """models/base_gattn.py:5-10,26-36,52-62,75-101 restated with stock torch ops: checked against numpy / sklearn."""
    import torch
    from sklearn.metrics import confusion_matrix, f1_score
    from han_b200.base_gattn import BaseGAttN
    rng = np.random.default_rng(3)
    M, C = 60, 4
    logits = torch.from_numpy(rng.normal(size=(M, C)))
    y = torch.from_numpy(rng.integers(0, C, size=M))
    cw = np.array([1.0, 2.0, 0.5, 3.0])
    lse = np.log(np.exp(logits.numpy()).sum(1))
    xent = lse - logits.numpy()[np.arange(M), y.numpy()]
    assert np.isclose(float(BaseGAttN.loss(logits, y, C, cw)), (xent * cw[y.numpy()]).mean())
    lg, lb = BaseGAttN.preshape(logits.reshape(1, M, C), y.reshape(1, M), C)
    assert lg.shape == (M, C) and lb.shape == (M,)
    cm = BaseGAttN.confmat(logits, y).numpy()
    assert np.array_equal(cm, confusion_matrix(y.numpy(), logits.numpy().argmax(1), labels=list(range(cm.shape[0]))))
    # multi-label helpers
    Y = torch.from_numpy((rng.random((M, C)) < 0.4).astype(np.int64))
    mask = torch.from_numpy((rng.random(M) < 0.5).astype(np.float32))
    x = logits.numpy()
    bce = (np.maximum(x, 0) - x * Y.numpy() + np.log1p(np.exp(-np.abs(x)))).mean(1)
    m = mask.numpy() / mask.numpy().mean()
    assert np.isclose(float(BaseGAttN.masked_sigmoid_cross_entropy(logits, Y, mask)), (bce * m).mean())
    sel = mask.numpy() > 0
    pred = (1 / (1 + np.exp(-x)) > 0.5).astype(int)
    assert np.isclose(float(BaseGAttN.micro_f1(logits, Y, mask)), f1_score(Y.numpy()[sel], pred[sel], average="micro"), atol=1e-6)
