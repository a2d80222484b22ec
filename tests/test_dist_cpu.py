"""Host-side logic of the multi-GPU path on CPU: the row partition, the merge of the per-rank
transposed slices into the by-source structure (vs scipy), and, under a real world_size-2 gloo
group, the edge exchange arithmetic, the sharded masked loss and the gradient all-reduce."""
import os

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as td
import torch.multiprocessing as mp

from han_b200 import dist as hd


def _random_csr(rng, n, deg):
    m = rng.random((n, n)) < deg / n
    np.fill_diagonal(m, True)
    return sp.csr_matrix(m.astype(np.int8))


def _shard_exchange_inputs(m, W, n_pad):
    """What rank r would send after transposing its destination-row shard: per source j the local
    destination ids (made global), source-major."""
    n = m.shape[0]
    sends = []
    for r in range(W):
        lo, hi = min(n, r * n_pad), min(n, (r + 1) * n_pad)
        t = m[lo:hi].T.tocsr()
        t.sort_indices()
        sends.append((np.diff(t.indptr).astype(np.int64), (t.indices + lo).astype(np.int32)))
    return sends


@pytest.mark.parametrize("n,W", [(50, 2), (37, 4), (64, 8), (10, 3)])
def test_merge_source_segments_matches_scipy(n, W):
    rng = np.random.default_rng(n * W)
    m = _random_csr(rng, n, 5.0)
    n_pad = -(-n // W)
    sends = _shard_exchange_inputs(m, W, n_pad)
    mt = m.T.tocsr()
    mt.sort_indices()
    for s in range(W):
        lo, hi = min(n, s * n_pad), min(n, (s + 1) * n_pad)
        n_loc = hi - lo
        counts = torch.stack([torch.from_numpy(sends[r][0][lo:hi]) for r in range(W)])
        segs = []
        for r in range(W):
            deg, idx = sends[r]
            off = np.concatenate([[0], np.cumsum(deg)])
            segs.append(torch.from_numpy(idx[off[lo]:off[hi]]))
        if n_loc == 0:
            continue
        indptr, indices = hd.merge_source_segments(counts, segs)
        ref = mt[lo:hi]
        assert np.array_equal(indptr.numpy(), ref.indptr.astype(np.int64))
        assert np.array_equal(indices.numpy(), ref.indices.astype(np.int32))


def test_row_range_covers_everything():
    for N, W in ((2_000_000, 8), (3025, 2), (7, 4), (736_389, 8)):
        spans = [hd.RowShard(r, W, torch.device("cpu")).row_range(N) for r in range(W)]
        assert spans[0][0] == 0 and spans[-1][1] == N
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        n_pad = -(-N // W)
        assert all(lo == min(N, r * n_pad) for r, (lo, _) in enumerate(spans))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    shard = hd.RowShard.init_process_group()
    try:
        assert shard.device.type == "cpu" and shard.world == world
        # (1) the edge exchange of RowShard._exchange, on host tensors: counts via all_to_all, then merge
        rng = np.random.default_rng(5)
        n = 41
        m = _random_csr(rng, n, 6.0)
        n_pad = -(-n // world)
        lo, hi = shard.row_range(n)
        t = m[lo:hi].T.tocsr(); t.sort_indices()
        deg_pad = torch.zeros(world * n_pad, dtype=torch.int64)
        deg_pad[:n] = torch.from_numpy(np.diff(t.indptr).astype(np.int64))
        recv_counts = torch.empty_like(deg_pad)
        td.all_to_all_single(recv_counts, deg_pad)
        recv_counts = recv_counts.view(world, n_pad)
        send_split = [int(x) for x in deg_pad.view(world, n_pad).sum(1)]
        recv_split = [int(x) for x in recv_counts.sum(1)]
        send = torch.from_numpy((t.indices + lo).astype(np.int32))
        recv = torch.empty(sum(recv_split), dtype=torch.int32)
        td.all_to_all_single(recv, send, output_split_sizes=recv_split, input_split_sizes=send_split)
        indptr, indices = hd.merge_source_segments(recv_counts[:, :hi - lo].contiguous(), list(torch.split(recv, recv_split)))
        ref = m.T.tocsr(); ref.sort_indices(); ref = ref[lo:hi]
        assert np.array_equal(indptr.numpy(), ref.indptr) and np.array_equal(indices.numpy(), ref.indices)

        # (2) sharded masked loss + gradient all-reduce == single-process value
        torch.manual_seed(0)
        N, C = 30, 3
        lin = torch.nn.Linear(4, C)
        X = torch.randn(N, 4)
        y = torch.nn.functional.one_hot(torch.randint(0, C, (N,)), C).float()
        mask = (torch.rand(N) < 0.5).float()

        class FakeTrain:
            def l2_loss(self):
                return 0.0005 * sum((p * p).sum() for p in lin.parameters())
        from han_b200.base_gattn import BaseGAttN
        ref_loss = BaseGAttN.masked_softmax_cross_entropy(lin(X), y, mask) + FakeTrain().l2_loss()
        ref_grads = torch.autograd.grad(ref_loss, list(lin.parameters()))
        lo, hi = shard.row_range(N)
        loss = shard.masked_loss(lin(X[lo:hi]), y[lo:hi], mask[lo:hi], FakeTrain())
        lin.zero_grad()
        loss.backward()
        shard.all_reduce_grads(lin)
        tot = shard.all_reduce_sum(loss.detach().clone().reshape(1))
        assert torch.allclose(tot[0], ref_loss, atol=1e-6)
        for p, g in zip(lin.parameters(), ref_grads):
            assert torch.allclose(p.grad, g, atol=1e-6)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        shard.shutdown()


def test_gloo_world2_exchange_loss_and_grads():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(msg == "ok" for _, msg in res), res


# ---- (meta-path x row-block) tile sharding: han_b200/tiles.py ------------------------------------------
def test_tile_layout_covers_every_tile_and_every_row_once():
    from han_b200.tiles import tile_layout, tile_rows
    for W, P, N in ((8, 4, 2_000_000), (4, 4, 3025), (2, 4, 97), (8, 2, 736_389), (2, 2, 10), (1, 3, 7)):
        owners = {}
        sem = []
        for r in range(W):
            Hn, h, Wz, member, paths = tile_layout(r, W, P)
            (a_lo, a_hi), (s_lo, s_hi), n_hpad, n_sub = tile_rows(N, Hn, h, Wz, member)
            assert Hn * Wz == W and len(paths) * Wz == P and a_lo <= s_lo <= s_hi <= a_hi
            for p in paths:
                assert (p, h) not in owners
                owners[(p, h)] = (a_lo, a_hi)
            sem.append((s_lo, s_hi))
        assert len(owners) == P * Hn
        for p in range(P):                               # the row blocks of one meta-path tile [0, N)
            spans = sorted(owners[(p, h)] for h in range(Hn))
            assert spans[0][0] == 0 and spans[-1][1] == N and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert sum(hi - lo for lo, hi in sem) == N and len({s for s in sem if s[1] > s[0]}) == sum(1 for s in sem if s[1] > s[0])
    with pytest.raises(ValueError):
        tile_layout(0, 3, 4)


def _tile_worker(rank, world, port, P, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from han_b200.tiles import TileShard
    tile = TileShard.init_process_group(P)
    try:
        N, D = 23, 6
        g = torch.Generator().manual_seed(11)
        Z_full = torch.randn(N, P, D, generator=g)
        Wt = torch.randn(N, P, D, generator=g)
        tile.bind([], N)
        (a_lo, a_hi), (s_lo, s_hi) = tile.attn_rows, tile.sem_rows
        Z_mine = Z_full[a_lo:a_hi][:, tile.paths, :].clone().requires_grad_(True)
        out = tile.exchange_Z(Z_mine)
        assert out.shape == (s_hi - s_lo, P, D) and torch.equal(out, Z_full[s_lo:s_hi])
        (out * Wt[s_lo:s_hi]).sum().backward()
        assert torch.equal(Z_mine.grad, Wt[a_lo:a_hi][:, tile.paths, :])
        # gradients of variables this rank never touched count as zero in the all-reduce
        lin_a, lin_b = torch.nn.Linear(3, 2), torch.nn.Linear(3, 2)
        with torch.no_grad():
            for p_ in list(lin_a.parameters()) + list(lin_b.parameters()):
                p_.fill_(0.5)
        mod = torch.nn.ModuleList([lin_a, lin_b])
        (lin_a if rank % 2 == 0 else lin_b)(torch.ones(1, 3)).sum().backward()
        tile.all_reduce_grads(mod)
        n_even = (world + 1) // 2
        assert torch.allclose(lin_a.weight.grad, torch.full((2, 3), float(n_even)))
        assert torch.allclose(lin_b.weight.grad, torch.full((2, 3), float(world - n_even)))
        q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        if td.is_initialized():
            td.destroy_process_group()


@pytest.mark.parametrize("world,P", [(2, 4), (2, 2), (4, 2)])
def test_gloo_tile_exchange_of_meta_path_embeddings(world, P):
    """world 2 / P 4: two meta-paths per rank, one row block.  world 4 / P 2: Hn = 2 row blocks per meta-path,
    the Z exchange runs inside the two ranks that share a row block (sub-groups)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29800 + (os.getpid() % 150) + 7 * world + P
    procs = [ctx.Process(target=_tile_worker, args=(r, world, port, P, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(msg == "ok" for _, msg in res), res
