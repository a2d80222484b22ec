"""Seeded synthetic inputs shaped like BASELINE.json's configs.

The reference's datasets are not in its checkout (.MISSING_LARGE_BLOBS), so every config is
synthetic.  Small configs (ACM / DBLP / IMDB-shaped) are generated with numpy on the host and
exist as dense masks too (so the dense reference path can run on them); the large ones
(2M-node, OGB-MAG-scale) are generated on the device directly as CSR.
Generation is rank-independent: every rank derives the same global graph from the seed.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np


@dataclass
class SmallConfig:
    name: str
    N: int
    F: int
    C: int
    metapaths: List[str]
    X: np.ndarray                    # (N,F) float32
    masks: List[np.ndarray]          # P boolean (N,N), self-loops included
    labels: np.ndarray               # (N,C) one-hot float32
    train_mask: np.ndarray           # (N,) bool
    val_mask: np.ndarray
    test_mask: np.ndarray

    @property
    def P(self):
        return len(self.masks)

    def adjs(self) -> List[np.ndarray]:
        """What the reference driver hands to adj_to_bias: (1,N,N) float64 'metapath - I'
        (ex_acm3025.py:61,110)."""
        eye = np.eye(self.N)
        return [(m.astype(np.float64) - eye)[None] for m in self.masks]

    def csr(self) -> List[Tuple[np.ndarray, np.ndarray]]:
        out = []
        for m in self.masks:
            rows, cols = np.nonzero(m)
            indptr = np.zeros(self.N + 1, dtype=np.int64)
            np.add.at(indptr, rows + 1, 1)
            out.append((np.cumsum(indptr).astype(np.int64), cols.astype(np.int32)))
        return out

    def n_edges(self) -> int:
        return int(sum(int(m.sum()) for m in self.masks))


def _sym_sparse(rng, n, mean_deg):
    """Symmetric random mask with self-loops and about ``mean_deg`` entries per row."""
    m = np.zeros((n, n), dtype=bool)
    n_pairs = int(max(0, (mean_deg - 1.0)) * n / 2)
    a = rng.integers(0, n, size=n_pairs)
    b = rng.integers(0, n, size=n_pairs)
    m[a, b] = True
    m[b, a] = True
    np.fill_diagonal(m, True)
    return m


def _cliques(rng, n, sizes_weights, memberships=1):
    """Union of cliques: node -> `memberships` groups drawn with probabilities `sizes_weights`."""
    g = len(sizes_weights)
    pw = np.asarray(sizes_weights, dtype=np.float64)
    pw = pw / pw.sum()
    member = np.zeros((n, g), dtype=bool)
    for _ in range(memberships):
        member[np.arange(n), rng.choice(g, size=n, p=pw)] = True
    m = (member.astype(np.float32) @ member.astype(np.float32).T) > 0
    np.fill_diagonal(m, True)
    return m


def _sym_dense(rng, n, density):
    u = rng.random((n, n)) < density / 2 * 1.0
    m = u | u.T
    np.fill_diagonal(m, True)
    return m


def _labels_masks(rng, n, c, n_train, n_val):
    y = rng.integers(0, c, size=n)
    onehot = np.zeros((n, c), dtype=np.float32)
    onehot[np.arange(n), y] = 1.0
    perm = rng.permutation(n)
    tr, va, te = (np.zeros(n, dtype=bool) for _ in range(3))
    tr[perm[:n_train]] = True
    va[perm[n_train:n_train + n_val]] = True
    te[perm[n_train + n_val:]] = True
    return onehot, tr, va, te


def acm_like(seed: int = 1000, scale: float = 1.0) -> SmallConfig:
    """cfg1: 3025 papers, 1870-d binary features, PAP (sparse, mean degree ~9.7) + PLP (56 skewed
    subject cliques, ~24% dense), 3 classes, 600/300/2125 split."""
    rng = np.random.default_rng(seed)
    n, f = int(3025 * scale), max(8, int(1870 * scale))
    X = (rng.random((n, f)) < 0.02).astype(np.float32)
    pap = _sym_sparse(rng, n, 9.7)
    plp = _cliques(rng, n, [r ** -1.5 for r in range(1, 57)], memberships=1)
    y, tr, va, te = _labels_masks(rng, n, 3, int(600 * scale), int(300 * scale))
    return SmallConfig("acm3025", n, f, 3, ["PAP", "PLP"], X, [pap, plp], y, tr, va, te)


def dblp_like(seed: int = 2000, scale: float = 1.0) -> SmallConfig:
    """cfg2: 4057 authors, 334-d binary features, APA (mean degree ~2.7), APCPA (20 conference
    cliques, ~30% dense), APTPA (~41% dense), 4 classes, 800/400/2857 split."""
    rng = np.random.default_rng(seed)
    n, f = int(4057 * scale), max(8, int(334 * scale))
    X = (rng.random((n, f)) < 0.05).astype(np.float32)
    apa = _sym_sparse(rng, n, 2.7)
    apcpa = _cliques(rng, n, [1.0 / (1 + 0.15 * r) for r in range(20)], memberships=3)
    aptpa = _sym_dense(rng, n, 0.41)
    y, tr, va, te = _labels_masks(rng, n, 4, int(800 * scale), int(400 * scale))
    return SmallConfig("dblp", n, f, 4, ["APA", "APCPA", "APTPA"], X, [apa, apcpa, aptpa], y, tr, va, te)


def imdb_like(seed: int = 3000, scale: float = 1.0) -> SmallConfig:
    """cfg3: 4780 movies, 1232-d binary features, MAM (mean degree ~19), MDM (~3.9), 3 classes,
    300/300/2687 split."""
    rng = np.random.default_rng(seed)
    n, f = int(4780 * scale), max(8, int(1232 * scale))
    X = (rng.random((n, f)) < 0.02).astype(np.float32)
    mam = _sym_sparse(rng, n, 19.0)
    mdm = _sym_sparse(rng, n, 3.9)
    y, tr, va, te = _labels_masks(rng, n, 3, int(300 * scale), int(300 * scale))
    return SmallConfig("imdb", n, f, 3, ["MAM", "MDM"], X, [mam, mdm], y, tr, va, te)


def tiny(seed: int = 7, n: int = 96, f: int = 40, p: int = 2, c: int = 3, deg: float = 6.0,
         binary: bool = False) -> SmallConfig:
    """Small random config for unit tests (real-valued features unless ``binary``)."""
    rng = np.random.default_rng(seed)
    X = (rng.random((n, f)) < 0.2).astype(np.float32) if binary else rng.normal(size=(n, f)).astype(np.float32)
    masks = []
    for i in range(p):
        m = rng.random((n, n)) < (deg - 1) / n
        np.fill_diagonal(m, True)
        masks.append(m)
    y, tr, va, te = _labels_masks(rng, n, c, n // 4, n // 8)
    return SmallConfig(f"tiny{n}", n, f, c, [f"MP{i}" for i in range(p)], X, masks, y, tr, va, te)


def planted(seed: int = 4000, n: int = 1200, f: int = 240, c: int = 3, homophily: float = 0.85,
            deg: float = 8.0, groups: int = 12, n_train: Optional[int] = None, n_val: Optional[int] = None) -> SmallConfig:
    """A LEARNABLE ACM-shaped heterograph for the training-loop example and its test (the shape-only
    configs above have uniform random labels): class-dependent bag-of-words features, a sparse
    homophilous meta-path ("PAP": a fraction ``homophily`` of the edges stay inside a class) and a
    clique meta-path ("PSP": ``groups`` subjects, each with a home class)."""
    rng = np.random.default_rng(seed)
    y = rng.integers(0, c, size=n)
    X = rng.random((n, f)) < 0.02
    w = f // (2 * c)
    for k in range(c):
        rows = np.nonzero(y == k)[0]
        X[np.ix_(rows, np.arange(k * w, (k + 1) * w))] |= rng.random((rows.size, w)) < 0.12
    X = X.astype(np.float32)
    by_class = [np.nonzero(y == k)[0] for k in range(c)]
    n_pairs = int(max(0.0, deg - 1.0) * n / 2)
    a = rng.integers(0, n, size=n_pairs)
    same = rng.random(n_pairs) < homophily
    b = rng.integers(0, n, size=n_pairs)
    for k in range(c):
        sel = np.nonzero(same & (y[a] == k))[0]
        b[sel] = by_class[k][rng.integers(0, by_class[k].size, size=sel.size)]
    pap = np.zeros((n, n), dtype=bool)
    pap[a, b] = True
    pap[b, a] = True
    np.fill_diagonal(pap, True)
    home = np.arange(groups) % c
    grp = rng.integers(0, groups, size=n)
    stay = rng.random(n) < homophily
    for k in range(c):
        sel = np.nonzero(stay & (y == k))[0]
        own = np.nonzero(home == k)[0]
        grp[sel] = own[rng.integers(0, own.size, size=sel.size)]
    psp = grp[:, None] == grp[None, :]
    np.fill_diagonal(psp, True)
    onehot = np.zeros((n, c), dtype=np.float32)
    onehot[np.arange(n), y] = 1.0
    n_train = n // 5 if n_train is None else n_train
    n_val = n // 10 if n_val is None else n_val
    perm = rng.permutation(n)
    tr, va, te = (np.zeros(n, dtype=bool) for _ in range(3))
    tr[perm[:n_train]] = True
    va[perm[n_train:n_train + n_val]] = True
    te[perm[n_train + n_val:]] = True
    return SmallConfig(f"planted{n}", n, f, c, ["PAP", "PSP"], X, [pap, psp], onehot, tr, va, te)


SMALL = {"acm": acm_like, "dblp": dblp_like, "imdb": imdb_like}



def _s64(x: int) -> int:
    """Python int -> the signed 64-bit value with the same bit pattern (torch int64 constants)."""
    x &= (1 << 64) - 1
    return x - (1 << 64) if x >= (1 << 63) else x


_K1, _K2, _K3 = _s64(0x9E3779B97F4A7C15), _s64(0xBF58476D1CE4E5B9), _s64(0x94D049BB133111EB)
_K4, _K5 = _s64(0xD6E8FEB86659FD93), _s64(0x632BE59BD9B4E019)


def _mix64(h):
    """splitmix64-style finaliser on an int64 tensor (wrapping arithmetic, arithmetic shifts)."""
    h = h ^ (h >> 30)
    h = h * _K2
    h = h ^ (h >> 27)
    h = h * _K3
    h = h ^ (h >> 31)
    return h & 0x7FFFFFFFFFFFFFFF

# ---------------------------------------------------------------------------------------------
# large configs: generated on the device
# ---------------------------------------------------------------------------------------------
@dataclass
class LargeSpec:
    name: str
    N: int
    F: int
    C: int
    P: int
    mean_degree: float
    powerlaw: bool = False
    seed: int = 4000


LARGE = {
    # cfg4: 2M nodes, 4 meta-paths, average degree 50, 256-d real-valued features
    "syn2m": LargeSpec("syn2m", 2_000_000, 256, 8, 4, 50.0, False, 4000),
    # cfg5: OGB-MAG scale: 736,389 papers, 2 power-law meta-paths, ~1e9 edges in total (the nominal mean degree is
    # before the cap at N and the de-duplication of a row's draws: 1170 gives ~1.0e9 distinct edges)
    "mag": LargeSpec("mag", 736_389, 128, 349, 2, 1170.0, True, 5000),
}


def device_random_csr(n_rows: int, n_cols: int, mean_degree: float, seed: int, device, row_lo: int = 0,
                      powerlaw: bool = False, chunk_rows: int = 1 << 18):
    """CSR (indptr int64, indices int32) for global rows [row_lo, row_lo + n_rows): each row has a
    self-loop plus random sources, sorted and de-duplicated.  Row r's content depends only on
    (seed, r), so any shard of rows can be generated independently and identically on every rank.
    Uniform: degree = mean_degree (before de-duplication).  Power-law: Zipf-like degrees
    (exponent 2.1) rescaled to the requested mean, capped at n_cols."""
    import torch
    counts, cols_out = [], []
    for c0 in range(0, n_rows, chunk_rows):
        c1 = min(n_rows, c0 + chunk_rows)
        nr = c1 - c0
        rows_g = torch.arange(row_lo + c0, row_lo + c1, device=device, dtype=torch.int64)
        if not powerlaw:
            d = int(round(mean_degree))
            # counter-based hash RNG: value depends on (seed, global row, slot) only
            slot = torch.arange(d - 1, device=device, dtype=torch.int64)
            src = _mix64(rows_g[:, None] * _K1 + slot[None, :] * _K2 + _s64(seed * _K3)) % n_cols
            allc = torch.cat([src, rows_g[:, None] % n_cols], dim=1)
            allc, _ = torch.sort(allc, dim=1)
            keep = torch.ones_like(allc, dtype=torch.bool)
            keep[:, 1:] = allc[:, 1:] != allc[:, :-1]
            counts.append(keep.sum(1))
            cols_out.append(allc[keep].to(torch.int32))
        else:
            # degrees: deg_r = clamp(scale * u^(-1/(a-1))), u from the row hash
            a = 2.1
            hr = _mix64(rows_g * _K1 + _s64(seed * _K4))
            u = ((hr & 0xFFFFFFFFFFFF).to(torch.float64) + 1.0) / float(1 << 48)
            raw = u.pow(-1.0 / (a - 1.0))
            # E[raw] for u~U(0,1): 1/(1 - 1/(a-1)) = 11 for a=2.1 (heavy tail; cap handles the rest)
            deg = (raw * (mean_degree / 11.0)).clamp(1, n_cols - 1).to(torch.int64)
            tot = int(deg.sum().item())
            seg = torch.repeat_interleave(torch.arange(nr, device=device), deg)
            start = torch.cumsum(deg, 0) - deg
            slot = torch.arange(tot, device=device, dtype=torch.int64) - start[seg]
            src = _mix64(rows_g[seg] * _K1 + slot * _K2 + _s64(seed * _K3)) % n_cols
            key = torch.cat([seg * n_cols + src, torch.arange(nr, device=device) * n_cols + (rows_g % n_cols)])
            key = torch.unique(key)  # sorted + de-duplicated; row-major order
            r = key // n_cols
            counts.append(torch.bincount(r, minlength=nr))
            cols_out.append((key - r * n_cols).to(torch.int32))
            del key, r, seg, slot, src
    cnt = torch.cat(counts)
    indptr = torch.zeros(n_rows + 1, dtype=torch.int64, device=device)
    indptr[1:] = torch.cumsum(cnt, 0)
    return indptr, torch.cat(cols_out)


def device_features(n_rows: int, F: int, seed: int, device, row_lo: int = 0, chunk_rows: int = 1 << 18):
    """X ~ N(0,1) fp32 for global rows [row_lo, row_lo+n_rows), reproducible per row block."""
    import torch
    out = torch.empty(n_rows, F, dtype=torch.float32, device=device)
    g = torch.Generator(device=device)
    # one generator state per fixed-size global chunk => independent of how rows are sharded
    first = row_lo // chunk_rows
    last = (row_lo + n_rows - 1) // chunk_rows
    for ch in range(first, last + 1):
        g.manual_seed(seed * 1_000_003 + ch)
        blk = torch.randn(chunk_rows, F, generator=g, device=device, dtype=torch.float32)
        lo = max(row_lo, ch * chunk_rows)
        hi = min(row_lo + n_rows, (ch + 1) * chunk_rows)
        out[lo - row_lo:hi - row_lo] = blk[lo - ch * chunk_rows:hi - ch * chunk_rows]
    return out


def device_labels(n_rows: int, C: int, seed: int, device, row_lo: int = 0):
    """(labels one-hot (n,C) fp32, train mask (n,) fp32 with ~20% ones), hash-based per global row."""
    import torch
    r = torch.arange(row_lo, row_lo + n_rows, device=device, dtype=torch.int64)
    h = _mix64(r * _K1 + _s64(seed * _K5))
    y = (h & 0x7FFFFFFF) % C
    onehot = torch.zeros(n_rows, C, dtype=torch.float32, device=device)
    onehot[torch.arange(n_rows, device=device), y] = 1.0
    mask = (((h >> 33) & 0xFFFF) % 5 == 0).to(torch.float32)
    return onehot, mask
