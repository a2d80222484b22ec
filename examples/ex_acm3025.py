"""Counterpart of the reference's training driver (ex_acm3025.py) on the B200 hot path.

    python examples/ex_acm3025.py                       # learnable synthetic ACM-shaped graph
    python examples/ex_acm3025.py --mat ACM3025.mat     # the real dataset, if you have it

Same hyper-parameters (:21-31), preprocessing (:57-61,110-118: ``adj = metapath - I`` -> ``adj_to_bias``),
training protocol (:161-247), test pass (:247-270) and embedding evaluation (:276-287).
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import han_b200 as hb                                   # noqa: E402
from han_b200 import jhyexps, synth, train              # noqa: E402


def sample_mask(idx, n):
    mask = np.zeros(n, dtype=bool)
    mask[np.asarray(idx).reshape(-1)] = True
    return mask


def load_mat(path):
    """ACM3025.mat layout (ex_acm3025.py:55-86): 'label' one-hot, 'feature', 'PAP', 'PLP', '*_idx'."""
    import scipy.io as sio
    data = sio.loadmat(path)
    y, X = data["label"].astype(np.float32), data["feature"].astype(np.float32)
    n = X.shape[0]
    adjs = [np.asarray(data[k], dtype=np.float64) - np.eye(n) for k in ("PAP", "PLP")]
    masks = [sample_mask(data[k], n) for k in ("train_idx", "val_idx", "test_idx")]
    return adjs, [X] * len(adjs), y, masks


def load_synthetic(n, seed):
    cfg = synth.planted(seed=seed, n=n)
    return [a[0] for a in cfg.adjs()], [cfg.X] * cfg.P, cfg.labels, [cfg.train_mask, cfg.val_mask, cfg.test_mask]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mat", default=None, help="path to ACM3025.mat (default: synthetic planted graph)")
    ap.add_argument("--nodes", type=int, default=3025)
    ap.add_argument("--epochs", type=int, default=200)
    ap.add_argument("--patience", type=int, default=100)
    ap.add_argument("--seed", type=int, default=4000)
    ap.add_argument("--checkpoint", default=None)
    ap.add_argument("--no-cuda-graph", dest="cuda_graph", action="store_false")
    args = ap.parse_args()

    adjs, feas, y, (tr, va, te) = load_mat(args.mat) if args.mat else load_synthetic(args.nodes, args.seed)
    n, c = y.shape
    print(f"Dataset: {'acm' if args.mat else 'synthetic-planted'}  nodes={n} features={feas[0].shape[1]} classes={c}")
    t0 = time.perf_counter()
    biases = [hb.process.adj_to_bias(a[None], [n], nhood=1) for a in adjs]       # :118, device CSR
    torch.cuda.synchronize()
    print(f"adj_to_bias: {len(biases)} meta-paths, edges={[g.nnz for g in biases]}, {time.perf_counter() - t0:.3f} s")
    splits = [np.where(m[:, None], y, 0.0).astype(np.float32) for m in (tr, va, te)]   # :72-77
    torch.manual_seed(args.seed)
    t0 = time.perf_counter()
    res = train.fit(feas, biases, *splits, tr, va, te, nb_epochs=args.epochs, patience=args.patience,
                    checkpt_file=args.checkpoint, cuda_graph=args.cuda_graph, log_every=10)
    dt = time.perf_counter() - t0
    print(f"{res.epochs_run} epochs in {dt:.2f} s ({1e3 * dt / res.epochs_run:.2f} ms/epoch incl. validation)")
    print("start knn, kmean.....")
    xx = res.final_embed.cpu().numpy()[te]
    yy = y[te]
    jhyexps.my_KNN(xx, yy, seed=0)
    jhyexps.my_Kmeans(xx, yy, k=c, seed=0)


if __name__ == "__main__":
    main()
