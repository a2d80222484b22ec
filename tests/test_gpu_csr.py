"""K-0 parity: the device CSR builder is bit-exact against the oracle's dense adj_to_bias
(np.nonzero(bias == 0) order), including the reference's edge cases, and the transposed structure
is a correct, deterministic permutation."""
import numpy as np
import pytest
import torch

from han_b200 import synth
from oracle import han_oracle as O

pytestmark = pytest.mark.gpu


def _csr_of_bias(bias):
    return O.bias_to_csr(bias)


def _assert_graph_equals(g, indptr, indices):
    ip, ix = g.to_host()
    assert ip.dtype == np.int64 and ix.dtype == np.int32
    assert np.array_equal(ip, indptr), "indptr differs"
    assert np.array_equal(ix, indices), "indices differ"


@pytest.mark.parametrize("n", [1, 31, 32, 33, 257, 1000])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_adj_to_csr_bit_exact(n, dtype):
    import han_b200 as hb
    rng = np.random.default_rng(n)
    adj = (rng.random((1, n, n)) < 0.05).astype(dtype) * rng.integers(1, 4, size=(1, n, n)).astype(dtype)
    adj[0] -= np.eye(n, dtype=dtype) * (rng.random(n) < 0.5)   # 'PAP - I' with some zero diagonals
    if n > 3:
        adj[0, 1, 2] = -3.0                                     # negative entry: stays masked
    ref = _csr_of_bias(O.adj_to_bias(adj.astype(np.float64), [n], 1))
    g = hb.process.adj_to_bias(adj, [n], nhood=1)
    assert g.shape == (1, n, n)
    _assert_graph_equals(g, *ref)
    # the same mask fed as a reference bias matrix (what ex_acm3025.py:180-181 feeds, cast to fp32)
    bias32 = O.adj_to_bias(adj.astype(np.float64), [n], 1).astype(np.float32)
    _assert_graph_equals(hb.MetaPathGraph.from_dense_bias(bias32), *ref)


def test_nhood2_bit_exact():
    import han_b200 as hb
    rng = np.random.default_rng(5)
    n = 200
    adj = (rng.random((1, n, n)) < 0.02).astype(np.float64)
    ref = _csr_of_bias(O.adj_to_bias(adj, [n], 2))
    _assert_graph_equals(hb.process.adj_to_bias(adj, [n], nhood=2), *ref)


def test_bias_with_unsupported_values_is_rejected():
    import han_b200 as hb
    b = np.full((1, 8, 8), -1e9, dtype=np.float32)
    b[0, np.arange(8), np.arange(8)] = 0
    b[0, 2, 3] = -0.5   # a genuine additive bias: not a mask
    with pytest.raises(ValueError):
        hb.MetaPathGraph.from_dense_bias(b)
    with pytest.raises(ValueError):
        hb.process.adj_to_bias(np.zeros((1, 8, 8)), [5])   # sizes[g] != N


@pytest.mark.parametrize("name", ["acm", "imdb"])
def test_config_shaped_graphs(name):
    import han_b200 as hb
    cfg = synth.SMALL[name](scale=0.35)
    for adj, (indptr, indices) in zip(cfg.adjs(), cfg.csr()):
        g = hb.process.adj_to_bias(adj, [cfg.N], nhood=1)
        _assert_graph_equals(g, indptr, indices)


def _check_transpose(g):
    import scipy.sparse as sp
    indptr, indices = g.to_host()
    n_r, n_c = g.n_rows, g.n_cols
    t0 = g.transpose()                         # keys only (nothing but edge weights needs the permutation)
    assert t0.perm is None
    tp0, ti0 = t0.to_host()
    t = g.transpose(with_perm=True)            # rebuilt with the permutation
    tp, ti = t.to_host()
    assert np.array_equal(tp0, tp) and np.array_equal(ti0, ti)
    perm = t.perm.cpu().numpy()
    m = sp.csr_matrix((np.arange(1, g.nnz + 1), indices, indptr), shape=(n_r, n_c))
    mt = m.T.tocsr()
    mt.sort_indices()
    assert np.array_equal(tp, mt.indptr.astype(np.int64))
    assert np.array_equal(ti, mt.indices.astype(np.int32))          # rows ascending within each column
    assert np.array_equal(perm, (mt.data - 1).astype(np.int32))     # edge positions in the CSR
    assert np.array_equal(np.sort(perm), np.arange(g.nnz))


def test_transpose_small_and_long_segments():
    import han_b200 as hb
    rng = np.random.default_rng(9)
    n = 6000
    m = rng.random((n, n)) < 0.002
    np.fill_diagonal(m, True)
    m[:, 17] = True                  # column of length 6000 (> 4096: global-memory sort path)
    m[:, 18] = rng.random(n) < 0.4   # ~2400 (shared-memory CTA sort path)
    m[:, 19] = rng.random(n) < 0.03  # ~180 (just above the warp path)
    g = hb.MetaPathGraph.from_dense_adj(m.astype(np.float32) - np.eye(n, dtype=np.float32))
    _check_transpose(g)
    _check_transpose(g.transpose())  # rectangular-safe and involutive structure


def test_transpose_rectangular_shard_and_empty():
    import han_b200 as hb
    cfg = synth.tiny(seed=4, n=300, f=4, p=1, deg=9.0)
    g = hb.process.adj_to_bias(cfg.adjs()[0], [cfg.N])
    shard = g.row_slice(100, 220)
    assert shard.n_rows == 120 and shard.n_cols == 300 and shard.row_offset == 100
    _check_transpose(shard)
    indptr = np.zeros(11, dtype=np.int64)
    empty = hb.MetaPathGraph.from_csr(indptr, np.zeros(0, dtype=np.int32), n_cols=10)
    t = empty.transpose()
    assert t.nnz == 0 and np.array_equal(t.to_host()[0], np.zeros(11, dtype=np.int64))


def test_from_csr_sort_and_duplicates():
    import han_b200 as hb
    rng = np.random.default_rng(3)
    n = 500
    rows = []
    for i in range(n):
        k = int(rng.integers(1, 300)) if i != 7 else 5000
        rows.append(rng.choice(20000, size=k, replace=False).astype(np.int32))
    indptr = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int64)
    indices = np.concatenate(rows)
    g = hb.MetaPathGraph.from_csr(indptr, indices, n_cols=20000, sort=True)
    _, ix = g.to_host()
    for i in range(n):
        assert np.array_equal(ix[indptr[i]:indptr[i + 1]], np.sort(rows[i]))
    bad = indices.copy()
    bad[indptr[3] + 1] = bad[indptr[3]]
    with pytest.raises(ValueError):
        hb.MetaPathGraph.from_csr(indptr, bad, n_cols=20000, sort=True)


def test_device_generator_is_shard_independent():
    dev = torch.device("cuda")
    ip, ix = synth.device_random_csr(5000, 100000, 50, 4000, dev)
    ip2, ix2 = synth.device_random_csr(1500, 100000, 50, 4000, dev, row_lo=2000)
    assert torch.equal(ix[ip[2000]:ip[3500]], ix2)
    d = (ip[1:] - ip[:-1])
    assert d.max().item() <= 50 and d.min().item() >= 45
    srt = torch.stack([ix[ip[r]:ip[r + 1]].diff().min() for r in range(0, 5000, 500)])
    assert (srt > 0).all()          # ascending, no duplicates
