"""Multi-GPU sharding of the hot path: one process per GPU, destination rows of every meta-path CSR
partitioned in contiguous equal blocks, NCCL over NVLink for the exchange steps.

The reference is single-process (ex_acm3025.py:163); this layer is new and its contract is "same
numbers as the single-device run" (SURVEY.md section 8(e)).

Per meta-path and step:
  forward : all-gather of the projected node table T = S          (n_pad x D per rank; f2 is recomputed)
  backward: all-gather of the row records R = [dV | f1 | lse | delta] (n_pad x RS per rank); the
            by-source pass then runs on the edges whose SOURCE is local; df1 is row-local (the forward
            keeps a second aggregate), so nothing is reduced back across ranks
  once    : all-reduce of the parameter gradients and of the loss.
Collectives run on a side stream and are ordered with events, so the gather of meta-path g+1
overlaps the aggregation of meta-path g.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import torch
import torch.distributed as td

from . import _lib
from ._lib import call, ptr, query, stream_ptr
from .graph import MetaPathGraph


def merge_source_segments(counts: torch.Tensor, segments: Sequence[torch.Tensor]):
    """Builds the by-source structure of the edges whose source is local from what every rank sent.

    counts   : (W, n_src) int64, counts[r, j] = number of destinations on rank r of local source j
    segments : W tensors; segments[r] lists, source-major, the GLOBAL destination ids on rank r
    Returns (indptr int64[n_src+1], indices int32[nnz]): per source, rank 0's destinations, then rank
    1's, ... -- ascending because destination blocks are ordered by rank.  Pure index arithmetic
    (device-agnostic), used once per graph.
    """
    W, n_src = counts.shape
    total = counts.sum(0)
    indptr = torch.zeros(n_src + 1, dtype=torch.int64, device=counts.device)
    indptr[1:] = torch.cumsum(total, 0)
    nnz = int(indptr[-1].item())
    indices = torch.empty(nnz, dtype=torch.int32, device=counts.device)
    before = torch.cumsum(counts, 0) - counts            # (W, n_src): edges of lower ranks for source j
    for r in range(W):
        c = counts[r]
        n_r = int(c.sum().item())
        if n_r == 0:
            continue
        seg_start = torch.cumsum(c, 0) - c                # offset of source j inside segments[r]
        src = torch.repeat_interleave(torch.arange(n_src, device=counts.device), c)
        within = torch.arange(n_r, device=counts.device) - seg_start[src]
        indices[indptr[:-1][src] + before[r][src] + within] = segments[r].to(torch.int32)
    return indptr, indices


class _BackwardEdges:
    """Edges (i, j) with j local, by source (for the gather pass of the backward)."""

    def __init__(self, by_src: MetaPathGraph):
        self.by_src = by_src                # rows = local sources, indices = global destination ids


class SymmetricTables:
    """Full-size node table T [G][W*n_pad][TS] and record table R [G][W*n_pad][RS] in symmetric memory
    (torch.distributed._symmetric_memory): every rank holds a copy at the same offsets and one NVLS
    MULTICAST address maps all copies, so a producer kernel's ``multimem.st`` lands in every rank's
    table (NVSwitch replicates it) -- the exchange is fused into the producers (projection epilogue,
    backward prep) instead of being a separate collective.  ``fence_*`` is the cross-rank barrier that
    says "everyone's rows have landed" / "nobody reads the old contents any more"."""

    def __init__(self, shard: "RowShard", G: int, TS: int, RS: int):
        import torch.distributed._symmetric_memory as symm
        dev = shard.device
        n_all = shard.world * shard.n_pad
        self.G, self.TS, self.RS, self.n_all = G, TS, RS, n_all
        grp = shard.group if shard.group is not None else td.group.WORLD     # a sub-group in tile mode
        self.T = symm.empty(G * n_all * TS, dtype=torch.float32, device=dev)
        self.hT = symm.rendezvous(self.T, grp)
        self.R = symm.empty(G * n_all * RS, dtype=torch.float32, device=dev)
        self.hR = symm.rendezvous(self.R, grp)
        self.T_mc = int(self.hT.multicast_ptr)
        self.R_mc = int(self.hR.multicast_ptr)
        if shard.comm == "multicast" and (not self.T_mc or not self.R_mc):
            raise RuntimeError("symmetric memory has no multicast (NVLS) mapping on this system")
        self.T.zero_()
        self.R.zero_()
        self.Tv = self.T.view(G, n_all, TS)
        self.Rv = self.R.view(G, n_all, RS)
        self.shard = shard
        W = shard.world
        # peer-mapped views of every rank's tables (for copy-engine pulls) and one side stream per peer
        self.peers_T = [self.hT.get_buffer(r, (G, n_all, TS), torch.float32) for r in range(W)]
        self.peers_R = [self.hR.get_buffer(r, (G, n_all, RS), torch.float32) for r in range(W)]
        self.pull_streams = [torch.cuda.Stream(device=dev, priority=-1) for _ in range(W - 1)]

    def T_mc_row(self, g: int, row: int):
        import ctypes
        return ctypes.c_void_p(self.T_mc + ((g * self.n_all + row) * self.TS) * 4)

    def R_mc_row(self, g: int, row: int):
        import ctypes
        return ctypes.c_void_p(self.R_mc + ((g * self.n_all + row) * self.RS) * 4)

    def fence_T(self, channel: int):
        self.hT.barrier(channel=channel)

    def fence_R(self, channel: int):
        self.hR.barrier(channel=channel)

    def _pull(self, hdl, view, peers):
        """After the cross-rank fence (every rank's block is complete), fetch the W-1 remote row blocks
        with peer-to-peer cudaMemcpyAsync: copy engines only, no SMs, so the transfers run fully
        concurrently with the aggregation kernels.  One side stream per peer, meta-paths in order;
        indexing the result with g waits only for table g."""
        sh = self.shard
        W, n_pad = sh.world, sh.n_pad
        cur = torch.cuda.current_stream()
        hdl.barrier(channel=0)
        _lib.trace_mark("fence done")
        ready = torch.cuda.Event()
        ready.record(cur)
        events = [[] for _ in range(self.G)]
        # table by table: the W-1 copies of table g run concurrently (one per peer), and table g+1 starts
        # only when table g is complete, so the consumer of table g is never starved by later tables
        for g in range(self.G):
            for k in range(1, W):
                r = (sh.rank + k) % W
                st = self.pull_streams[k - 1]
                if g == 0:
                    st.wait_event(ready)
                else:
                    for ev in events[g - 1]:
                        st.wait_event(ev)
                with torch.cuda.stream(st):
                    view[g, r * n_pad:(r + 1) * n_pad].copy_(peers[r][g, r * n_pad:(r + 1) * n_pad], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(st)
                    events[g].append(ev)
                    _lib.trace_mark(f"pull g{g} peer{r} done")
        return _Pulled(view, events)

    # ---- "push" exchange: producers hand finished row chunks of their block to every peer right away ----------
    PUSH_CHUNKS = int(os.environ.get("HAN_PUSH_CHUNKS", "4"))
    PUSH_MIN_ROWS = int(os.environ.get("HAN_PUSH_MIN_ROWS", "4096"))     # tests lower it to chunk tiny graphs too

    def chunk_bounds(self, n: int):
        """Row chunks of this rank's block for the chunked producers (multiples of 128 rows: projection tiles)."""
        c = max(1, min(self.PUSH_CHUNKS, n // self.PUSH_MIN_ROWS))
        step = -(-(-(-n // c)) // 128) * 128
        return [(r0, min(n, r0 + step)) for r0 in range(0, n, step)]

    def _push(self, view, peers, g, r0, r1):
        """Copy rows [r0, r1) (table coordinates) of table g (None: all G) of MY copy into every peer's copy, on the
        push stream, ordered after the work queued so far on the current stream."""
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(cur)
        st = self.pull_streams[0]
        st.wait_event(ev)
        W = self.shard.world
        with torch.cuda.stream(st):
            for k in range(1, W):
                r = (self.shard.rank + k) % W
                for gg in (range(self.G) if g is None else (g,)):
                    peers[r][gg, r0:r1].copy_(view[gg, r0:r1], non_blocking=True)
        _lib.trace_mark(f"push rows {r0}:{r1} queued")

    def _pushed(self, hdl, view):
        """All pushes of every rank have landed: one cross-rank barrier on the push stream; the current stream waits."""
        st = self.pull_streams[0]
        with torch.cuda.stream(st):
            hdl.barrier(channel=0)
            done = torch.cuda.Event()
            done.record(st)
        torch.cuda.current_stream().wait_event(done)
        _lib.trace_mark("pushes landed")
        return view

    def push_T(self, r0: int, r1: int):
        self._push(self.Tv, self.peers_T, None, r0, r1)

    def pushed_T(self):
        return self._pushed(self.hT, self.Tv)

    def push_R(self, g: int, r0: int, r1: int):
        self._push(self.Rv, self.peers_R, g, r0, r1)

    def pushed_R(self):
        return self._pushed(self.hR, self.Rv)

    def exchange_T(self, fused_multicast: bool):
        if fused_multicast:
            self.fence_T(0)           # producers already wrote every rank's copy (multimem.st)
            return self.Tv
        return self._pull(self.hT, self.Tv, self.peers_T)

    def exchange_R(self, fused_multicast: bool):
        if fused_multicast:
            self.fence_R(0)
            return self.Rv
        return self._pull(self.hR, self.Rv, self.peers_R)


class _Pulled:
    def __init__(self, view, events):
        self.view, self.events = view, events

    def __getitem__(self, g: int) -> torch.Tensor:
        cur = torch.cuda.current_stream()
        for ev in self.events[g]:
            cur.wait_event(ev)
        _lib.trace_mark(f"table {g} ready")
        return self.view[g]


class RowShard:
    def __init__(self, rank: int, world: int, device: torch.device, group=None):
        self.rank, self.world, self.device, self.group = rank, world, device, group
        self.n_total = None
        self.n_pad = None
        self._tables = {}
        # exchange of the node tables / row records between ranks:
        #   "pull"      (default) symmetric memory + copy-engine peer copies, overlapped with the kernels
        #   "multicast" producers write every rank's copy through the NVLS multicast address (multimem.st)
        #   "nccl"      plain NCCL all-gathers on a high-priority side stream
        #   "push"      symmetric memory; the producers (projection, backward prep) run in row chunks and copy every
        #               finished chunk into the peers' tables while the next chunk is computed (tile sharding's default)
        self.comm = os.environ.get("HAN_DIST_COMM", "pull")
        self.use_multicast = self.comm != "nccl"
        self.comm_stream = torch.cuda.Stream(device=device, priority=-1) if device.type == "cuda" else None
        self._bwd = {}

    # ---- set-up ------------------------------------------------------------------------------
    @staticmethod
    def init_process_group() -> "RowShard":
        rank = int(os.environ["RANK"])
        world = int(os.environ["WORLD_SIZE"])
        local = int(os.environ.get("LOCAL_RANK", rank))
        if torch.cuda.is_available():
            dev = torch.device("cuda", local)
            torch.cuda.set_device(dev)
            # High-priority NCCL stream: the gather kernels fill every SM (3-4 CTAs each holding ~55-68 KB
            # of shared memory), so a default-priority collective launched next to them only starts when
            # the whole compute grid has drained -- i.e. no overlap at all.  With priority its CTAs are
            # placed as soon as compute CTAs retire (every ~100 us).
            opts = td.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            td.init_process_group("nccl", device_id=dev, pg_options=opts)
        else:
            dev = torch.device("cpu")
            td.init_process_group("gloo")
        return RowShard(rank, world, dev)

    def row_range(self, N: int):
        """Contiguous equal blocks of ceil(N/W) rows (the last block may be shorter)."""
        n_pad = -(-N // self.world)
        lo = min(N, self.rank * n_pad)
        return lo, min(N, lo + n_pad)

    def bind(self, graphs: Sequence[MetaPathGraph], N: int) -> None:
        """Once per graph set: exchange the transposed slices so every rank owns the by-source
        structure of the edges whose source it owns (df1 is row-local, so nothing else is needed)."""
        self.n_total = N
        self.n_pad = -(-N // self.world)
        for g in graphs:
            if id(g) in self._bwd:
                continue
            self._bwd[id(g)] = self._exchange(g)

    def _exchange(self, g: MetaPathGraph) -> _BackwardEdges:
        W, n_pad, N = self.world, self.n_pad, self.n_total
        dev = g.device
        lo, hi = self.row_range(N)
        gt = g.transpose()                                  # rows = ALL sources j, entries = local dest ids
        if W == 1:
            # the whole meta-path is local (a tile-sharded rank that owns whole meta-paths): the by-source view IS the
            # structure of the edges whose source is local -- no exchange, no merge, no host round trip
            return _BackwardEdges(gt)
        deg = (gt.indptr[1:] - gt.indptr[:-1])               # (N,)
        deg_pad = torch.zeros(W * n_pad, dtype=torch.int64, device=dev)
        deg_pad[:N] = deg
        recv_counts = torch.empty(W * n_pad, dtype=torch.int64, device=dev)
        td.all_to_all_single(recv_counts, deg_pad, group=self.group)     # counts[r, j_local]
        recv_counts = recv_counts.view(W, n_pad)
        send_split = deg_pad.view(W, n_pad).sum(1).tolist()
        recv_split = recv_counts.sum(1).tolist()
        send = (gt.indices.to(torch.int64) + lo).to(torch.int32)          # global destination ids
        recv = torch.empty(int(sum(recv_split)), dtype=torch.int32, device=dev)
        td.all_to_all_single(recv, send, output_split_sizes=[int(x) for x in recv_split],
                             input_split_sizes=[int(x) for x in send_split], group=self.group)
        n_loc = hi - lo
        segs = list(torch.split(recv, [int(x) for x in recv_split]))
        indptr, indices = merge_source_segments(recv_counts[:, :n_loc].contiguous(), segs)
        return _BackwardEdges(MetaPathGraph(indptr, indices, n_loc, W * n_pad))

    def symmetric_tables(self, G: int, K: int, H: int, slot: int = 0) -> Optional[SymmetricTables]:
        """The multicast-mapped tables for this plan shape (allocated and rendezvoused once; a
        collective call, so every rank must ask for the same shapes in the same order), or None when
        NVLS multicast is unavailable / HAN_DIST_COMM=nccl (then NCCL all-gathers are used)."""
        if not self.use_multicast or self.device.type != "cuda":
            return None
        # One table set per plan invocation that is alive between a forward and its backward (`slot`: group index
        # of the first layer, 1000*layer + meta-path for stacked layers): the backward reads T and R straight from
        # the tables, so two plans of the same shape must never share one.
        key = (G, K, H, self.n_pad, int(slot))
        if key not in self._tables:
            TS, RS = query("han_table_stride", K, H), query("han_record_stride", K, H)
            # No silent downgrade: a system without symmetric memory must be run with HAN_DIST_COMM=nccl explicitly.
            self._tables[key] = SymmetricTables(self, G, TS, RS)
        return self._tables[key]

    # ---- collectives ---------------------------------------------------------------------------
    def barrier(self):
        td.barrier(group=self.group)

    def all_reduce_sum(self, t: torch.Tensor) -> torch.Tensor:
        td.all_reduce(t, op=td.ReduceOp.SUM, group=self.group)
        return t

    def all_reduce_max(self, t: torch.Tensor) -> torch.Tensor:
        td.all_reduce(t, op=td.ReduceOp.MAX, group=self.group)
        return t

    def _pad_rows(self, t: torch.Tensor) -> torch.Tensor:
        """(.., n_loc, C) -> (.., n_pad, C); only the last rank ever pads."""
        n_loc = t.shape[-2]
        if n_loc == self.n_pad:
            return t
        out = torch.zeros(*t.shape[:-2], self.n_pad, t.shape[-1], dtype=t.dtype, device=t.device)
        out[..., :n_loc, :] = t
        return out

    def all_gather_rows(self, T: torch.Tensor) -> List[torch.Tensor]:
        """T (G, n_loc, C) -> list of G lazily-completed (W*n_pad, C) tables.  The gathers run on the
        side stream in meta-path order; ``wait(g)`` makes the current stream wait for table g."""
        G = T.shape[0]
        Tp = self._pad_rows(T)
        full = [torch.empty(self.world * self.n_pad, T.shape[-1], dtype=T.dtype, device=T.device) for _ in range(G)]
        return _Gathered(self, Tp, full)

    # ---- the sharded by-source backward of one meta-path ---------------------------------------
    def gather_records(self, R: torch.Tensor):
        """R (G, n_loc, RS) -> gathered handle (see all_gather_rows); issued for all meta-paths at once
        so that gather g+1 overlaps the by-source pass of g."""
        return self.all_gather_rows(R)

    def backward_edges(self, plan, g: int, T_local: torch.Tensor, a2: torch.Tensor, b2: torch.Tensor, R_full: torch.Tensor,
                       dS: torch.Tensor, df2: torch.Tensor) -> None:
        """Runs the by-source gather pass on the edges whose source is local (dS, df2 of the local source rows) against
        the records of ALL destination rows.  Nothing comes back across ranks: df1 is row-local (ops: prep kernel)."""
        be: _BackwardEdges = self._bwd[id(plan.graphs[g])]
        K, H = plan.K, plan.H
        n_loc = T_local.shape[0]
        bs = be.by_src
        row0 = self.row_range(self.n_total)[0]
        tv = bs.split_view()
        if tv is not None:      # heavy source rows (power-law meta-paths): virtual-row view + merge
            part = torch.empty((tv.n_slots, K, H + 2), dtype=torch.float32, device=T_local.device)
            call("han_attn_bwd_src_chunked_split", ptr(tv.indptr_v), ptr(bs.indices), ptr(tv.chunk_rows), tv.n_chunks, n_loc,
                 ptr(T_local), ptr(a2), ptr(b2), ptr(R_full), K, H, ptr(dS), ptr(df2), None, ptr(plan.seed),
                 1.0 - plan.coef_drop, 1.0 - plan.in_drop, plan.metapath_id(g), row0, ptr(tv.vmap), ptr(part), ptr(tv.heavy_rows), ptr(tv.heavy_ptr), tv.n_heavy,
                 stream_ptr())
        else:
            cr, n_chunks = bs.chunks()
            call("han_attn_bwd_src_chunked", ptr(bs.indptr), ptr(bs.indices), ptr(cr), n_chunks, n_loc, ptr(T_local),
                 ptr(a2), ptr(b2), ptr(R_full), K, H, ptr(dS), ptr(df2), None, ptr(plan.seed), 1.0 - plan.coef_drop,
                 1.0 - plan.in_drop, plan.metapath_id(g), row0, stream_ptr())

    # ---- loss / gradients ------------------------------------------------------------------------
    def masked_loss(self, logits, labels, mask, train_op):
        """This rank's share of masked CE (models/base_gattn.py:41-48 over the GLOBAL mask) plus
        1/W of the L2 term, so that the all-reduced gradients equal the single-process ones."""
        mask = mask.to(logits.dtype)
        mask_total = self.all_reduce_sum(mask.sum().reshape(1))          # stays on the device: no host sync
        if logits.is_cuda and logits.dtype == torch.float32:
            from . import ops
            ce = ops.masked_ce(logits, labels, mask, mask_total)
        else:
            labels = labels.to(logits.dtype)
            xent = -(labels * torch.log_softmax(logits, dim=-1)).sum(-1)
            ce = ((xent * mask).sum() / mask_total).squeeze(0)
        return ce if train_op is None else ce + train_op.l2_loss() / self.world

    def all_reduce_flat(self, flat: torch.Tensor) -> None:
        """Sum one flat gradient buffer over ranks (the fused optimizer's layout): a single collective."""
        _lib.trace_mark("all_reduce grads >")
        td.all_reduce(flat, op=td.ReduceOp.SUM, group=self.group)
        _lib.trace_mark("all_reduce grads <")

    def all_reduce_grads(self, module: torch.nn.Module) -> None:
        grads = [p.grad for p in module.parameters() if p.grad is not None]
        flat = torch.cat([g.reshape(-1) for g in grads])
        _lib.trace_mark("all_reduce grads >")
        td.all_reduce(flat, op=td.ReduceOp.SUM, group=self.group)
        _lib.trace_mark("all_reduce grads <")
        off = 0
        for g in grads:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n

    def shutdown(self):
        if td.is_initialized():
            td.destroy_process_group()


class _Gathered:
    """G all-gathers issued back to back on the side stream; indexing waits for that table only."""

    def __init__(self, shard: RowShard, Tp: torch.Tensor, full: List[torch.Tensor]):
        self.full = full
        self.events = []
        cur = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(cur)
        with torch.cuda.stream(shard.comm_stream):
            shard.comm_stream.wait_event(ready)
            for g in range(Tp.shape[0]):
                td.all_gather_into_tensor(full[g], Tp[g].contiguous(), group=shard.group)
                ev = torch.cuda.Event()
                ev.record(shard.comm_stream)
                self.events.append(ev)
                _lib.trace_mark(f"nccl gather {g} done")
        for t in full:
            t.record_stream(shard.comm_stream)
        Tp.record_stream(shard.comm_stream)

    def __getitem__(self, g: int) -> torch.Tensor:
        torch.cuda.current_stream().wait_event(self.events[g])
        _lib.trace_mark(f"table {g} ready")
        return self.full[g]
