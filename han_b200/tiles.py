"""(meta-path x row-block) tile sharding -- the low-traffic alternative to ``dist.RowShard``.

Row sharding (dist.py) gives every rank a block of destination rows of EVERY meta-path, so every rank
must fetch every other rank's node-table rows ``T`` and row records ``R`` of every meta-path: at 8 GPUs
on the 2M-node graph 4.46 GB per rank and step, and the step is bound by that exchange (DESIGN.md
section 8).  Here rank ``r = h*P + p`` owns ONE meta-path p and, when there are more ranks than
meta-paths, one of ``Hn = W/P`` row blocks h of it (with fewer ranks than meta-paths a rank owns
``P/W`` whole meta-paths):

  * node-level attention of meta-path p needs ``T_p`` / ``R_p`` only: they are exchanged inside the Hn
    ranks that share p (an ordinary ``RowShard`` on that sub-group; no exchange at all when Hn == 1);
  * the semantic layer needs all P meta-path embeddings of a node on one rank: ``exchange_Z`` re-shards
    ``Z`` from (meta-path, row block h) tiles to row sub-blocks with ONE all-to-all inside the ranks
    that share h (and the backward sends ``dZ`` the way back);
  * parameters are replicated; a rank's gradients for the meta-paths it does not own are zero and the
    usual all-reduce sums the rest.

Traffic per rank and step on the 2M graph at 8 GPUs: 256 + 352 MB inside the pair, 2 x 192 MB of
``Z`` / ``dZ`` = about 1.0 GB instead of 4.46 GB.  Equal-size tiles assume meta-paths of similar weight;
edge-balanced tile assignment is future work.

bench.py picks this partitioning whenever the rank and meta-path counts divide (``--partition auto``; ``row`` / ``tile``
force one); host arithmetic is covered by gloo tests (tests/test_dist_cpu.py), the whole step by tests/dist_check.py
(``HAN_DIST_PARTITION=tile``) and by the parity leg of every multi-GPU bench run.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as td

from . import _lib
from .dist import RowShard


def tile_layout(rank: int, world: int, P: int) -> Tuple[int, int, int, int, List[int]]:
    """-> (Hn row blocks per meta-path, h = this rank's row block, Wz = ranks sharing h, member = index
    inside those, paths = the meta-paths this rank owns)."""
    if world >= P:
        if world % P:
            raise ValueError(f"tile sharding needs the rank count ({world}) to be a multiple of the meta-paths ({P})")
        Hn, per = world // P, 1
        h, member = rank // P, rank % P
    else:
        if P % world:
            raise ValueError(f"tile sharding needs the meta-paths ({P}) to be a multiple of the rank count ({world})")
        Hn, per = 1, P // world
        h, member = 0, rank
    return Hn, h, world // Hn, member, list(range(member * per, (member + 1) * per))


def tile_rows(N: int, Hn: int, h: int, Wz: int, member: int) -> Tuple[Tuple[int, int], Tuple[int, int], int, int]:
    """-> ((attention rows lo, hi), (semantic rows lo, hi), n_hpad, n_sub): attention block h of
    ceil(N/Hn) rows, cut into Wz semantic sub-blocks of ceil(n_hpad/Wz) rows."""
    n_hpad = -(-N // Hn)
    a_lo = min(N, h * n_hpad)
    a_hi = min(N, a_lo + n_hpad)
    n_sub = -(-n_hpad // Wz)
    s_lo = min(a_hi, a_lo + member * n_sub)
    s_hi = min(a_hi, s_lo + n_sub)
    return (a_lo, a_hi), (s_lo, s_hi), n_hpad, n_sub


class ZBuffers:
    """Symmetric-memory landing zones of the two re-sharding steps between node-level attention (meta-path x
    row-block tiles) and the semantic layer (row sub-blocks, all meta-paths), inside the Wz ranks that share a row block:

      Zrecv  [n_sub][P][D]      this rank's semantic rows, ALL meta-paths: member m's K-B kernels store their meta-paths'
                                columns of these rows straight into it over NVLink (han_attn_fwd_chunked's `out2`), so the
                                all-to-all of Z costs no pass of its own and overlaps the rest of the gather;
      dZrecv [Wz*n_sub][per][D] the gradient of this rank's attention rows for its own meta-paths: every member pushes the
                                slice of its semantic backward that belongs here (peer-to-peer copies).

    One cross-rank barrier per direction says "everything has landed"."""

    def __init__(self, tile: "TileShard", D: int):
        import torch.distributed._symmetric_memory as symm
        dev = tile.device
        P, Wz, n_sub, per = tile.P, tile.Wz, tile.n_sub, len(tile.paths)
        grp = tile.z_group if tile.z_group is not None else td.group.WORLD
        self.D = D
        self.zr = symm.empty(n_sub * P * D, dtype=torch.float32, device=dev)
        self.hz = symm.rendezvous(self.zr, grp)
        self.dz = symm.empty(Wz * n_sub * per * D, dtype=torch.float32, device=dev)
        self.hd = symm.rendezvous(self.dz, grp)
        self.zr.zero_()
        self.dz.zero_()
        self.Zrecv = self.zr.view(n_sub, P, D)
        self.dZrecv = self.dz.view(Wz * n_sub, per, D)
        self.peer_Z = [self.hz.get_buffer(r, (n_sub, P, D), torch.float32) for r in range(Wz)]
        self.peer_dZ = [self.hd.get_buffer(r, (Wz * n_sub, per, D), torch.float32) for r in range(Wz)]

    def barrier_Z(self):
        self.hz.barrier(channel=0)

    def barrier_dZ(self):
        self.hd.barrier(channel=0)


class ZSink:
    """What ops.NodeAttentionFn needs to fuse the Z re-sharding into K-B: the row sub-blocks of this rank's attention
    rows and, per sub-block and local meta-path, the peer address its output rows also go to."""

    def __init__(self, tile: "TileShard", bufs: ZBuffers):
        self.tile, self.bufs = tile, bufs
        self.used = False

    def out2(self, g: int):
        """(device table of base pointers, rows per block, row stride in floats) for local meta-path g: block m of this
        rank's attention rows (n_sub rows) is member m's semantic rows, and its rows land in member m's Zrecv at local
        meta-path g's columns."""
        t, b = self.tile, self.bufs
        cache = b.__dict__.setdefault("_out2_tabs", {})
        if g not in cache:
            col = (t.member * len(t.paths) + g) * b.D
            ptrs = [b.peer_Z[m].data_ptr() + col * 4 for m in range(t.Wz)]
            cache[g] = torch.tensor(ptrs, dtype=torch.int64, device=t.device)
        return _lib.ptr(cache[g]), t.n_sub, t.P * b.D


class _ZExchange(torch.autograd.Function):
    """Z (rows of block h, this rank's meta-paths, D)  ->  (this rank's semantic rows, ALL P meta-paths, D).

    Forward: when K-B already stored its rows into the owners' Zrecv (``pushed``), only the barrier remains; otherwise
    (stacked layers, heavy-row graphs) the slices are pushed here with peer-to-peer copies.  Backward: the slices of dZ
    go to the ranks that own those meta-paths (peer-to-peer copies into their dZrecv), one barrier."""

    @staticmethod
    def forward(ctx, Z, tile: "TileShard", pushed: bool):
        n_h, per, D = Z.shape
        b = tile.zbuffers(D)
        if not pushed:
            for m in range(tile.Wz):
                r0, r1 = min(n_h, m * tile.n_sub), min(n_h, (m + 1) * tile.n_sub)
                if r1 > r0:
                    b.peer_Z[m][:r1 - r0, tile.member * per:(tile.member + 1) * per, :].copy_(Z[r0:r1], non_blocking=True)
        _lib.trace_mark("Z landed >")
        b.barrier_Z()
        _lib.trace_mark("Z landed <")
        ctx.tile, ctx.n_h, ctx.per = tile, n_h, per
        n_sem = tile.sem_rows[1] - tile.sem_rows[0]
        return b.Zrecv[:n_sem]

    @staticmethod
    def backward(ctx, dOut):
        tile, n_h, per = ctx.tile, ctx.n_h, ctx.per
        n_sem, P, D = dOut.shape
        b = tile.zbuffers(D)
        me = tile.member
        _lib.trace_mark("dZ push >")
        if tile._dz_routed:
            tile._dz_routed = False           # the semantic backward kernel stored its rows into the owners' dZrecv itself
        else:
            for m in range(tile.Wz):
                if n_sem:
                    b.peer_dZ[m][me * tile.n_sub:me * tile.n_sub + n_sem].copy_(dOut[:, m * per:(m + 1) * per, :],
                                                                                non_blocking=True)
        b.barrier_dZ()
        _lib.trace_mark("dZ push <")
        return b.dZrecv[:n_h], None, None


class _ZExchangeA2A(torch.autograd.Function):
    """The same re-sharding as ONE ``all_to_all_single`` each way: used with the gloo backend (CPU tests of the layout
    arithmetic) and with HAN_DIST_COMM=nccl."""

    @staticmethod
    def forward(ctx, Z, tile: "TileShard"):
        n_h, per, D = Z.shape
        Wz, n_sub = tile.Wz, tile.n_sub
        send = Z.new_zeros(Wz * n_sub, per, D)
        send[:n_h] = Z
        recv = torch.empty_like(send)
        _lib.trace_mark("all_to_all Z >")
        td.all_to_all_single(recv, send, group=tile.z_group)
        _lib.trace_mark("all_to_all Z <")
        ctx.tile, ctx.n_h = tile, n_h
        n_sem = tile.sem_rows[1] - tile.sem_rows[0]
        # recv[j] = member j's meta-paths for MY sub-block: (Wz, n_sub, per, D) -> (n_sub, Wz*per = P, D)
        return recv.view(Wz, n_sub, per, D).permute(1, 0, 2, 3).reshape(n_sub, Wz * per, D)[:n_sem].contiguous()

    @staticmethod
    def backward(ctx, dOut):
        tile, n_h = ctx.tile, ctx.n_h
        Wz, n_sub = tile.Wz, tile.n_sub
        n_sem, P, D = dOut.shape
        per = P // Wz
        pad = dOut.new_zeros(n_sub, P, D)
        pad[:n_sem] = dOut
        send = pad.view(n_sub, Wz, per, D).permute(1, 0, 2, 3).contiguous()
        recv = torch.empty_like(send)
        _lib.trace_mark("all_to_all dZ >")
        td.all_to_all_single(recv, send, group=tile.z_group)
        _lib.trace_mark("all_to_all dZ <")
        return recv.view(Wz * n_sub, per, D)[:n_h].contiguous(), None


class TileShard:
    def __init__(self, rank: int, world: int, P: int, device: torch.device):
        self.rank, self.world, self.P, self.device = rank, world, P, device
        self.Hn, self.h, self.Wz, self.member, self.paths = tile_layout(rank, world, P)
        self.z_group = None                      # WORLD when every rank shares the one row block
        self.attn: Optional[RowShard] = None     # exchange of T / R inside the ranks that share a meta-path
        if self.Hn > 1:
            # collective: every rank creates every group, in the same order
            zg = [td.new_group([h * P + p for p in range(P)]) for h in range(self.Hn)]
            ag = [td.new_group([h * P + p for h in range(self.Hn)]) for p in range(P)]
            self.z_group = zg[self.h]
            # T / R exchange inside the ranks that share a meta-path: symmetric-memory tables on that sub-group with
            # copy-engine pulls (HAN_DIST_COMM=pull, default) or NCCL all-gathers (HAN_DIST_COMM=nccl)
            self.attn = RowShard(self.h, self.Hn, device, group=ag[self.member])
            # default here: "push" (one table per rank: nothing to pipeline table by table, so the producers are chunked
            # and each chunk travels while the next is computed)
            import os
            self.attn.comm = os.environ.get("HAN_TILE_COMM", "push" if self.attn.comm != "nccl" else "nccl")
            self.attn.use_multicast = self.attn.comm != "nccl"
        self.n_total = None
        self.attn_rows = self.sem_rows = (0, 0)
        self.n_hpad = self.n_sub = 0
        self._zbufs = {}
        self._dz_routed = False
        import os
        # Z / dZ re-sharding through symmetric memory (K-B's fused stores + peer copies), or as an all-to-all collective
        self.fused_z = device.type == "cuda" and os.environ.get("HAN_DIST_COMM", "pull") != "nccl"

    @staticmethod
    def init_process_group(P: int) -> "TileShard":
        shard = RowShard.init_process_group()    # same backend / device set-up as row sharding
        return TileShard(shard.rank, shard.world, P, shard.device)

    # ---- set-up ------------------------------------------------------------------------------
    def rows(self, N: int):
        """-> (attention rows (lo, hi), semantic rows (lo, hi)) of this rank for an N-node graph."""
        a, s, _, _ = tile_rows(N, self.Hn, self.h, self.Wz, self.member)
        return a, s

    def bind(self, graphs: Sequence, N: int) -> None:
        """Once per graph set (graphs = this rank's meta-paths, rows = its attention block)."""
        self.n_total = N
        self.attn_rows, self.sem_rows, self.n_hpad, self.n_sub = tile_rows(N, self.Hn, self.h, self.Wz, self.member)
        if self.attn is not None:
            self.attn.bind(graphs, N)
        else:
            for g in graphs:
                g.transpose()

    def reset(self) -> None:
        if self.attn is not None:
            self.attn._bwd = {}

    # ---- the one exchange of the forward / backward ------------------------------------------------
    def zbuffers(self, D: int) -> ZBuffers:
        """Symmetric landing zones for this graph size (collective: every rank of the row block calls it in step)."""
        key = (self.n_sub, D)
        if key not in self._zbufs:
            self._zbufs[key] = ZBuffers(self, D)
        return self._zbufs[key]

    def z_sink(self, D: int) -> Optional[ZSink]:
        """Hand this to the LAST attention layer's plans: their K-B kernels then store every output row into the
        owner's semantic input as well (the all-to-all fused into the kernel's epilogue).  None when the exchange
        runs as an NCCL / gloo all-to-all."""
        return ZSink(self, self.zbuffers(D)) if self.fused_z else None

    def gather_features(self, host_rows: torch.Tensor, stream: Optional[torch.cuda.Stream] = None,
                        slot: int = 0) -> torch.Tensor:
        """Host-fed runs: every rank uploads only ITS semantic rows of the feature matrix (1/W of it, pinned host memory
        -> its slice of a symmetric buffer) and hands the slice to the other ranks of its row block over NVLink
        (peer-to-peer copies, ~600 GB/s against the 20-55 GB/s of the host link).  Returns the features of this rank's
        attention rows [n_h][F].  Stream-ordered on ``stream`` (default: current); the caller waits on that stream.
        ``slot`` selects one of several symmetric buffers, so that a pipelined caller can stage step k+1 while step k
        still reads its features."""
        import torch.distributed._symmetric_memory as symm
        F = host_rows.shape[1]
        n_sem = self.sem_rows[1] - self.sem_rows[0]
        assert host_rows.shape[0] == n_sem
        key = ("X", self.n_sub, F, slot)
        if key not in self._zbufs:
            grp = self.z_group if self.z_group is not None else td.group.WORLD
            t = symm.empty(self.Wz * self.n_sub * F, dtype=torch.float32, device=self.device)
            h = symm.rendezvous(t, grp)
            peers = [h.get_buffer(r, (self.Wz * self.n_sub, F), torch.float32) for r in range(self.Wz)]
            self._zbufs[key] = (t.view(self.Wz * self.n_sub, F), h, peers)
        X, h, peers = self._zbufs[key]
        st = stream if stream is not None else torch.cuda.current_stream()
        lo = self.member * self.n_sub
        with torch.cuda.stream(st):
            mine = X[lo:lo + n_sem]
            mine.copy_(host_rows, non_blocking=True)
            for k in range(1, self.Wz):
                m = (self.member + k) % self.Wz
                peers[m][lo:lo + n_sem].copy_(mine, non_blocking=True)
            h.barrier(channel=0)
        return X[:self.attn_rows[1] - self.attn_rows[0]]

    def dz_route(self, n_sem: int, P: int, D: int):
        """For the semantic backward kernel: (device table of P base pointers, row stride in floats) so that the gradient
        row of (semantic row i, meta-path p) lands in the owner of p's dZrecv -- or None when the exchange runs as a
        collective.  Marks the pending exchange as already routed."""
        if not self.fused_z or P != self.P or n_sem != self.sem_rows[1] - self.sem_rows[0]:
            return None
        b = self.zbuffers(D)
        if not hasattr(b, "_dz_tab"):
            per = len(self.paths)
            ptrs = [b.peer_dZ[p // per].data_ptr() + ((self.member * self.n_sub) * per + p % per) * D * 4 for p in range(P)]
            b._dz_tab = torch.tensor(ptrs, dtype=torch.int64, device=self.device)
        self._dz_routed = True
        return _lib.ptr(b._dz_tab), len(self.paths) * D

    def exchange_Z(self, Z: torch.Tensor, pushed: bool = False) -> torch.Tensor:
        if self.fused_z:
            return _ZExchange.apply(Z, self, pushed)
        return _ZExchangeA2A.apply(Z, self)

    # ---- collectives over all ranks ------------------------------------------------------------------
    def barrier(self):
        td.barrier()

    def all_reduce_sum(self, t: torch.Tensor) -> torch.Tensor:
        td.all_reduce(t, op=td.ReduceOp.SUM)
        return t

    def all_reduce_max(self, t: torch.Tensor) -> torch.Tensor:
        td.all_reduce(t, op=td.ReduceOp.MAX)
        return t

    def masked_loss(self, logits, labels, mask, train_op):
        """This rank's share of masked CE over the GLOBAL mask (models/base_gattn.py:41-48) plus 1/W of the L2
        term; logits / labels / mask are this rank's SEMANTIC rows."""
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

    def all_reduce_grads(self, module: torch.nn.Module) -> None:
        """Variables of the meta-paths this rank does not own have no gradient here: they count as zero."""
        params = list(module.parameters())
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
        _lib.trace_mark("all_reduce grads >")
        td.all_reduce(flat, op=td.ReduceOp.SUM)
        _lib.trace_mark("all_reduce grads <")
        off = 0
        for p in params:
            n = p.numel()
            g = flat[off:off + n].view_as(p)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += n

    def all_reduce_flat(self, flat: torch.Tensor) -> None:
        td.all_reduce(flat, op=td.ReduceOp.SUM)
