"""Embedding-quality protocol of the paper's Tables 3/4 as the reference driver runs it on the test
embeddings (ex_acm3025.py:276-287 -> jhyexp.py:20-86): k-NN classification at several training
ratios (macro/micro F1, averaged over reshuffles) and k-means clustering (NMI / ARI, averaged over
restarts).  Host-side evaluation glue on scikit-learn -- not part of the GPU hot path; the functions
keep the reference's names and defaults, print the same summary lines, and also RETURN the scores.
"""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

import numpy as np


def _labels_1d(y) -> np.ndarray:
    y = np.asarray(y)
    return y.argmax(axis=1) if y.ndim > 1 else y


def my_KNN(x, y, k: int = 5, split_list: Sequence[float] = (0.2, 0.4, 0.6, 0.8), time: int = 10,
           show_train: bool = True, shuffle: bool = True, seed=None) -> Dict[float, Tuple[float, float]]:
    """jhyexp.py:20-53.  Returns {split: (macro_f1, micro_f1)}; the first ``split`` fraction of a fresh
    permutation trains the classifier, the rest is scored."""
    from sklearn.metrics import f1_score
    from sklearn.neighbors import KNeighborsClassifier
    rng = np.random.default_rng(seed)
    x = np.squeeze(np.asarray(x))
    y = _labels_1d(y)
    out = {}
    for ratio in split_list:
        cut = int(x.shape[0] * ratio)
        macro, micro = [], []
        for _ in range(max(1, time)):
            order = rng.permutation(x.shape[0]) if shuffle else np.arange(x.shape[0])
            xs, ys = x[order], y[order]
            pred = KNeighborsClassifier(n_neighbors=k).fit(xs[:cut], ys[:cut]).predict(xs[cut:])
            macro.append(f1_score(ys[cut:], pred, average="macro"))
            micro.append(f1_score(ys[cut:], pred, average="micro"))
        out[ratio] = (float(np.mean(macro)), float(np.mean(micro)))
        print("KNN({}avg, split:{}, k={}) f1_macro: {:.4f}, f1_micro: {:.4f}".format(time, ratio, k, *out[ratio]))
    return out


def my_Kmeans(x, y, k: int = 4, time: int = 10, return_NMI: bool = False, seed=None):
    """jhyexp.py:55-86.  Mean NMI / ARI of ``time`` k-means runs against the labels."""
    from sklearn.cluster import KMeans
    from sklearn.metrics import adjusted_rand_score, normalized_mutual_info_score
    x = np.squeeze(np.asarray(x))
    y = _labels_1d(y)
    rs = np.random.RandomState(seed)
    nmi, ari = [], []
    for _ in range(max(1, time)):
        pred = KMeans(n_clusters=k, n_init=10, random_state=rs.randint(2 ** 31 - 1)).fit_predict(x)
        nmi.append(normalized_mutual_info_score(y, pred))
        ari.append(adjusted_rand_score(y, pred))
    score, s2 = float(np.mean(nmi)), float(np.mean(ari))
    print("NMI ({} avg): {:.4f} , ARI ({}avg): {:.4f}".format(time, score, time, s2))
    if return_NMI:
        return score, s2
