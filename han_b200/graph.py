"""Meta-path graph handle: the CSR (and lazily its transpose) that replaces the reference's dense
``bias_mat`` (utils/process.py:14-25, fed as (1,N,N) fp32 every step at ex_acm3025.py:180-181).

Built on the device by the K-0 kernels; bit-exact against ``np.nonzero(bias == 0)``.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import call, ptr, query, stream_ptr


def _dev(device=None) -> torch.device:
    if device is not None:
        return torch.device(device)
    if not torch.cuda.is_available():
        raise _lib.HanError("han_b200 needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _as_device_tensor(a, device, dtype=None) -> torch.Tensor:
    if isinstance(a, np.ndarray):
        a = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None and a.dtype != dtype:
        a = a.to(dtype)
    return a.to(device, non_blocking=True).contiguous()


# rows longer than this are cut into segments for the edge-stream kernels (MetaPathGraph.split_view)
SPLIT_ROW_EDGES = int(os.environ.get("HAN_SPLIT_ROW_EDGES", "4096"))


def split_layout(indptr: torch.Tensor, S: int):
    """Index arithmetic of ``MetaPathGraph.split_view`` (any device): cut every row of the CSR ``indptr`` with
    more than S edges into ceil(deg/S) segments of at most S edges.  Returns
    (indptr_v int64[n_v+1], vptr int64[n+1], vmap int32[n_v][2] = (real row, partial slot or -1),
    heavy_rows int32[n_h], heavy_ptr int32[n_h+1], n_slots)."""
    dev = indptr.device
    n = indptr.numel() - 1
    deg = indptr[1:] - indptr[:-1]
    nseg = torch.clamp((deg + (S - 1)) // S, min=1)
    vptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    vptr[1:] = torch.cumsum(nseg, 0)
    vrow = torch.repeat_interleave(torch.arange(n, device=dev, dtype=torch.int64), nseg)
    n_v = int(vrow.numel())
    seg = torch.arange(n_v, device=dev, dtype=torch.int64) - vptr[vrow]
    indptr_v = torch.empty(n_v + 1, dtype=torch.int64, device=dev)
    indptr_v[:-1] = indptr[vrow] + seg * S
    indptr_v[-1] = indptr[-1]
    heavy = nseg > 1
    hv_v = heavy[vrow]
    slot = torch.cumsum(hv_v.to(torch.int64), 0) - 1
    vmap = torch.stack([vrow, torch.where(hv_v, slot, torch.full_like(slot, -1))], 1).to(torch.int32).contiguous()
    heavy_rows = torch.nonzero(heavy).reshape(-1).to(torch.int32)
    heavy_ptr = torch.zeros(heavy_rows.numel() + 1, dtype=torch.int32, device=dev)
    heavy_ptr[1:] = torch.cumsum(nseg[heavy], 0).to(torch.int32)
    return indptr_v, vptr, vmap, heavy_rows, heavy_ptr, int(hv_v.sum().item())


@dataclass
class SplitView:
    indptr_v: torch.Tensor      # int64 [n_v+1]: the CSR offsets with the cut points inserted
    vptr: torch.Tensor          # int64 [n_rows+1]: virtual rows of each real row
    vmap: torch.Tensor          # int32 [n_v][2]: (real row, partial slot or -1)
    heavy_rows: torch.Tensor    # int32 [n_heavy]
    heavy_ptr: torch.Tensor     # int32 [n_heavy+1]: partial slots of each cut row
    chunk_rows: torch.Tensor    # int32 [n_chunks+1] over the virtual rows
    n_chunks: int
    n_v: int
    n_heavy: int
    n_slots: int


class MetaPathGraph:
    """CSR of one meta-path mask.  Rows = destination nodes i (softmax rows of
    utils/layers.py:27), columns = source nodes j; columns ascending within a row.

    ``row_offset``: global id of local row 0 when this is a destination-row shard.
    """

    def __init__(self, indptr: torch.Tensor, indices: torch.Tensor, n_rows: int, n_cols: int,
                 nnz: Optional[int] = None, row_offset: int = 0):
        assert indptr.dtype == torch.int64 and indices.dtype == torch.int32
        assert indptr.is_cuda and indices.is_cuda
        self.indptr = indptr
        self.indices = indices
        self.n_rows = int(n_rows)
        self.n_cols = int(n_cols)
        self.nnz = int(indices.numel() if nnz is None else nnz)
        self.row_offset = int(row_offset)
        self._t: Optional["MetaPathGraph"] = None
        self.perm: Optional[torch.Tensor] = None  # set on the transposed view
        self._empty_rows: Optional[bool] = None
        # set when the arrays were produced on another stream (host->device staging, transposition on a
        # side stream): consumers call wait_ready() before their first kernel that reads this graph
        self.ready: Optional[torch.cuda.Event] = None
        # sp_attn_head only (utils/layers.py:95-96): the stored adjacency values w_ij (fp32, CSR order) that scale
        # the logits, l_ij = w_ij (f1_i + f2_j); None = the 0/1 adjacency every other operator uses
        self.edge_weight: Optional[torch.Tensor] = None

    def wait_ready(self) -> "MetaPathGraph":
        if self.ready is not None:
            torch.cuda.current_stream(self.device).wait_event(self.ready)
        return self

    # the reference hands around a (1,N,N) array; keep that visible for call-site compatibility
    @property
    def shape(self):
        return (1, self.n_rows, self.n_cols)

    @property
    def device(self):
        return self.indptr.device

    def has_empty_rows(self) -> bool:
        """Rows without any edge (the dense path turns those into uniform 1/N rows)."""
        if self._empty_rows is None:
            d = self.indptr[1:] - self.indptr[:-1]
            self._empty_rows = bool((d == 0).any().item())
        return self._empty_rows

    # ---- constructors -----------------------------------------------------------------------
    @staticmethod
    def _from_dense(dense, kind: int, device=None) -> "MetaPathGraph":
        device = _dev(device)
        if isinstance(dense, np.ndarray):
            dense = torch.from_numpy(np.ascontiguousarray(dense))
        if dense.dim() == 3:
            if dense.shape[0] != 1:
                raise ValueError("one graph per call: expected (1,N,N) or (N,N)")
            dense = dense[0]
        if dense.dim() != 2 or dense.shape[0] != dense.shape[1]:
            raise ValueError(f"expected a square matrix, got {tuple(dense.shape)}")
        if dense.dtype not in (torch.float32, torch.float64):
            dense = dense.to(torch.float64)
        dense = dense.to(device).contiguous()
        dt = _lib.F32 if dense.dtype == torch.float32 else _lib.F64
        n = dense.shape[0]
        with torch.cuda.device(device):
            counts = torch.empty(n, dtype=torch.int32, device=device)
            bad = torch.zeros(1, dtype=torch.int32, device=device)
            call("han_dense_row_counts", ptr(dense), dt, kind, n, n, ptr(counts), ptr(bad), stream_ptr())
            indptr = torch.empty(n + 1, dtype=torch.int64, device=device)
            ws_bytes = query("han_scan_workspace_bytes", n)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
            call("han_scan_counts", ptr(counts), n, ptr(indptr), ptr(ws), ws_bytes, stream_ptr())
            nnz, nbad = int(indptr[-1].item()), int(bad.item())  # build-time sync (not in the step)
            if kind == _lib.DENSE_BIAS and nbad:
                raise ValueError(
                    f"bias_mat has {nbad} entries that are neither 0 nor <= -1e8: only the reference's "
                    "mask biases (utils/process.py:25) are supported")
            indices = torch.empty(max(nnz, 1), dtype=torch.int32, device=device)[:nnz]
            if nnz:
                call("han_dense_fill_indices", ptr(dense), dt, kind, n, n, ptr(indptr), ptr(indices),
                     stream_ptr())
        return MetaPathGraph(indptr, indices, n, n, nnz)

    @staticmethod
    def from_dense_adj(adj, nhood: int = 1, device=None) -> "MetaPathGraph":
        """``adj_to_bias(adj, [N], nhood)`` semantics (utils/process.py:14-25): edge (i,j) iff
        ((adj + I)^nhood)_ij > 0.  nhood == 1 is evaluated in one pass by the K-0 kernels; for
        nhood > 1 the fp64 matrix power of the reference (:19-20) is taken on the device first."""
        if nhood < 1:
            # reference: zero loop iterations leave mt = I -> only self-loops
            raise ValueError("nhood must be >= 1")
        if nhood == 1:
            return MetaPathGraph._from_dense(adj, _lib.DENSE_ADJ, device)
        device = _dev(device)
        a = _as_device_tensor(adj, device, torch.float64)
        if a.dim() == 3:
            a = a[0]
        eye = torch.eye(a.shape[0], dtype=torch.float64, device=device)
        mt = eye.clone()
        for _ in range(nhood):
            mt = mt @ (a + eye)
        return MetaPathGraph._from_dense(mt, _lib.DENSE_POSITIVE, device)  # :21-24: mt > 0

    @staticmethod
    def from_dense_bias(bias, device=None) -> "MetaPathGraph":
        """From a reference bias matrix (values 0 / -1e9, (1,N,N) or (N,N))."""
        return MetaPathGraph._from_dense(bias, _lib.DENSE_BIAS, device)

    @staticmethod
    def from_csr(indptr, indices, n_cols: Optional[int] = None, device=None, sort: bool = False,
                 row_offset: int = 0, stream: Optional[torch.cuda.Stream] = None) -> "MetaPathGraph":
        """From host or device CSR arrays.  ``sort=True`` sorts columns within rows on the device
        and rejects duplicate edges.  ``stream``: stage the (pinned) host arrays on that stream instead of
        the current one -- the handle carries a ``ready`` event and the kernels wait for it, so the copy
        overlaps whatever the compute stream is doing."""
        device = _dev(device)
        if stream is not None:
            if sort:
                raise ValueError("sort=True needs the arrays on the current stream")
            nnz_host = int(indices.numel() if isinstance(indices, torch.Tensor) else len(indices))
            host_deg = None
            if isinstance(indptr, torch.Tensor) and not indptr.is_cuda and indptr.numel() > 1:
                host_deg = indptr[1:] - indptr[:-1]      # host arrays: the degree facts cost no device round trip
            with torch.cuda.stream(stream):
                indptr = _as_device_tensor(indptr, device, torch.int64)
                indices = _as_device_tensor(indices, device, torch.int32)
                ev = torch.cuda.Event()
                ev.record(stream)
            g = MetaPathGraph(indptr, indices, indptr.numel() - 1, indptr.numel() - 1 if n_cols is None else n_cols,
                              nnz_host, row_offset=row_offset)
            g.ready = ev
            if host_deg is not None:
                g._empty_rows = bool((host_deg == 0).any())
                g._max_deg = int(host_deg.max())
            return g
        indptr = _as_device_tensor(indptr, device, torch.int64)
        indices = _as_device_tensor(indices, device, torch.int32)
        n_rows = indptr.numel() - 1
        if n_cols is None:
            n_cols = n_rows
        g = MetaPathGraph(indptr, indices, n_rows, n_cols, row_offset=row_offset)
        if sort and g.nnz:
            with torch.cuda.device(device):
                scratch = torch.zeros(n_rows + 64, dtype=torch.int32, device=device)
                dup = torch.zeros(1, dtype=torch.int32, device=device)
                call("han_csr_sort_rows", n_rows, ptr(indptr), ptr(indices), None, ptr(scratch), ptr(dup),
                     stream_ptr())
                if int(dup.item()):
                    raise ValueError(f"{int(dup.item())} duplicate edges in CSR input")
        return g

    # ---- derived structures -----------------------------------------------------------------
    def transpose(self, stream: Optional[torch.cuda.Stream] = None, with_perm: Optional[bool] = None) -> "MetaPathGraph":
        """By-source view: row j lists the destinations i of edges (i,j), ascending.  ``with_perm`` also builds
        ``perm`` = position of each transposed edge in this CSR (default: only when the graph carries edge weights --
        nothing else needs it since the backward keeps no per-edge array).  Built once, cached.  ``stream``: build it
        on that stream (after this graph is ready there); the view then carries its own ``ready`` event."""
        if with_perm is None:
            with_perm = self.edge_weight is not None
        if self._t is not None and with_perm and self._t.perm is None:
            self._t = None                  # rebuild with the permutation
        if self._t is None and stream is not None:
            with torch.cuda.stream(stream):
                self.wait_ready()
                t = self.transpose(with_perm=with_perm)
                # the view's host-side facts (maximum degree -> virtual rows or not) and its work-item table are taken
                # here, so that their one device->host read waits for THIS stream only, not for the compute stream
                if t.split_view() is None:
                    t.chunks()
                t.ready = torch.cuda.Event()
                t.ready.record(stream)
            return t
        if self._t is None:
            device = self.device
            self.wait_ready()
            with torch.cuda.device(device):
                t_indptr = torch.empty(self.n_cols + 1, dtype=torch.int64, device=device)
                t_indices = torch.empty(max(self.nnz, 1), dtype=torch.int32, device=device)[:self.nnz]
                perm = torch.empty(max(self.nnz, 1), dtype=torch.int32, device=device)[:self.nnz] if with_perm else None
                ws_bytes = query("han_transpose_workspace_bytes", self.n_rows, self.n_cols, self.nnz)
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
                call("han_csr_transpose", self.n_rows, self.n_cols, self.nnz, ptr(self.indptr),
                     ptr(self.indices), ptr(t_indptr), ptr(t_indices), ptr(perm), ptr(ws), ws_bytes,
                     stream_ptr())
            t = MetaPathGraph(t_indptr, t_indices, self.n_cols, self.n_rows, self.nnz)
            t.perm = perm
            self._t = t
        return self._t

    def chunks(self):
        """(chunk_rows int32[n_chunks+1], n_chunks): whole-row work items of ~2048 edges for the
        chunked edge-stream kernels.  Built once per graph, cached."""
        if getattr(self, "_chunks", None) is None:
            self.wait_ready()
            n_chunks = int(query("han_csr_num_chunks", self.nnz))
            with torch.cuda.device(self.device):
                cr = torch.empty(n_chunks + 1, dtype=torch.int32, device=self.device)
                call("han_csr_chunk_rows", ptr(self.indptr), self.n_rows, self.nnz, ptr(cr), stream_ptr())
            self._chunks = (cr, n_chunks)
        return self._chunks

    def split_view(self):
        """Virtual-row view for graphs with heavy rows (power-law meta-paths): every row with more than
        ``SPLIT_ROW_EDGES`` edges is cut into segments of at most that many, so that no single warp of the
        edge-stream kernels is left with a 10^5-edge row.  Returns None when no row is that long (the
        common case: nothing changes), else a ``SplitView``.  Built once per graph, cached; the one
        device->host read (the maximum degree) happens here, outside any captured step."""
        if getattr(self, "_split", False) is False:
            self._split = None
            self.wait_ready()
            S = SPLIT_ROW_EDGES
            max_deg = getattr(self, "_max_deg", None)
            if max_deg is None and self.nnz > 0:
                max_deg = int((self.indptr[1:] - self.indptr[:-1]).max().item())
            if self.nnz > 0 and max_deg > S:
                dev, n = self.device, self.n_rows
                with torch.cuda.device(dev):
                    indptr_v, vptr, vmap, heavy_rows, heavy_ptr, n_slots = split_layout(self.indptr, S)
                    n_v = int(indptr_v.numel()) - 1
                    n_chunks = int(query("han_csr_num_chunks", self.nnz))
                    cr = torch.empty(n_chunks + 1, dtype=torch.int32, device=dev)
                    call("han_csr_chunk_rows", ptr(indptr_v), n_v, self.nnz, ptr(cr), stream_ptr())
                self._split = SplitView(indptr_v, vptr, vmap, heavy_rows, heavy_ptr, cr, n_chunks, n_v,
                                        int(heavy_rows.numel()), n_slots)
        return self._split

    def edge_weight_t(self) -> Optional[torch.Tensor]:
        """The edge weights in transposed-edge order (``edge_weight[perm]``) for the by-source backward pass;
        None for a 0/1 adjacency.  Built once per graph, cached."""
        if self.edge_weight is None:
            return None
        if getattr(self, "_ew_t", None) is None:
            self._ew_t = self.edge_weight[self.transpose(with_perm=True).perm.long()].contiguous()
        return self._ew_t

    def row_slice(self, lo: int, hi: int) -> "MetaPathGraph":
        """Destination-row shard [lo, hi) (column ids stay global)."""
        base = int(self.indptr[lo].item())
        end = int(self.indptr[hi].item())
        indptr = (self.indptr[lo:hi + 1] - base).contiguous()
        indices = self.indices[base:end].contiguous()
        return MetaPathGraph(indptr, indices, hi - lo, self.n_cols, end - base, row_offset=lo)

    # ---- host views (tests / debugging) -------------------------------------------------------
    def to_host(self):
        return self.indptr.cpu().numpy(), self.indices.cpu().numpy()

    def to_dense_bias(self, dtype=np.float32) -> np.ndarray:
        """The (1,N,N) bias matrix the reference would feed: 0 on edges, -1e9 elsewhere."""
        indptr, indices = self.to_host()
        b = np.full((self.n_rows, self.n_cols), -1e9, dtype=dtype)
        rows = np.repeat(np.arange(self.n_rows), np.diff(indptr))
        b[rows, indices] = 0.0
        return b[None]
