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
