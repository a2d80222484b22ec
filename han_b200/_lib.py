"""ctypes binding of libhan_sm100.so (include/han_b200.h).  No fallback: if the CUDA library is
missing or a call fails, this raises."""
from __future__ import annotations

import ctypes
from ctypes import c_char_p, c_int, c_int32, c_int64, c_size_t, c_void_p
from pathlib import Path

import torch

_LIB_PATH = Path(__file__).resolve().parent / "libhan_sm100.so"

ACT_IDENTITY, ACT_ELU = 0, 1
SEM_REFERENCE, SEM_PAPER = 0, 1
DENSE_ADJ, DENSE_BIAS, DENSE_POSITIVE = 0, 1, 2
F32, F64 = 0, 1

P = c_void_p
I64 = c_int64
I = c_int
SZ = c_size_t
FL = ctypes.c_float

# name -> (restype, argtypes); mirrors include/han_b200.h one to one
_SIGNATURES = {
    "han_version": (c_int, []),
    "han_last_error": (c_char_p, []),
    "han_attn_shape_supported": (c_int, [I, I]),
    "han_table_stride": (c_int, [I, I]),
    "han_record_stride": (c_int, [I, I]),
    "han_dense_row_counts": (c_int, [P, I, I, I64, I64, P, P, P]),
    "han_scan_workspace_bytes": (SZ, [I64]),
    "han_scan_counts": (c_int, [P, I64, P, P, SZ, P]),
    "han_dense_fill_indices": (c_int, [P, I, I, I64, I64, P, P, P]),
    "han_transpose_workspace_bytes": (SZ, [I64, I64, I64]),
    "han_csr_transpose": (c_int, [I64, I64, I64, P, P, P, P, P, P, SZ, P]),
    "han_csr_sort_rows": (c_int, [I64, P, P, P, P, P, P]),
    "han_project_fwd": (c_int, [P, I64, I64, I64, P, I, I, I, P, P, P, P, I, P]),
    "han_project_tc_workspace_bytes": (SZ, [I64, I, I, I]),
    "han_project_fwd_tc": (c_int, [P, I64, I64, I64, P, I, I, I, P, P, P, P, P, I64, I64, I64, I, P, SZ, P]),
    "han_multicast_copy": (c_int, [P, P, I64, P]),
    "han_project_bwd_workspace_bytes": (SZ, [I64, I64, I, I]),
    "han_project_bwd": (c_int, [P, I64, I64, I64, P, I, I, P, P, SZ, I, P]),
    "han_project_bwd_tc_workspace_bytes": (SZ, [I64, I64, I]),
    "han_project_bwd_tc": (c_int, [P, I64, I64, I64, P, I, P, P, SZ, I, P]),
    "han_attn_coefs": (c_int, [P, P, I64, P, P, P, P, I, I, P, P, P]),
    "han_csr_chunk_edges": (c_int64, [I64]),
    "han_csr_num_chunks": (c_int64, [I64]),
    "han_csr_chunk_rows": (c_int, [P, I64, I64, P, P]),
    "han_attn_fwd_chunked": (c_int, [P, P, P, I64, I64, P, P, P, P, P, I, I, I, P, I64, P, P, P, P, I64, P, I64, I64, P, P, P,
                                     FL, FL, I, I64, P]),
    "han_attn_bwd_src_chunked": (c_int, [P, P, P, I64, I64, P, P, P, P, I, I, P, P, P, P, FL, FL, I, I64, P]),
    "han_attn_fwd_chunked_split": (c_int, [P, P, P, I64, I64, P, P, P, P, P, I, I, I, P, I64, P, P, P, P, I64, P, I64, I64, P, P,
                                           P, FL, FL, I, I64, P, P, P, P, I, P]),
    "han_attn_bwd_src_chunked_split": (c_int, [P, P, P, I64, I64, P, P, P, P, I, I, P, P, P, P, FL, FL, I, I64, P, P, P, P, I,
                                               P]),
    "han_project_fwd_drop": (c_int, [P, I64, I64, I64, P, I64, I, I, I, P, P, P, P, P, FL, I, I64, P]),
    "han_project_bwd_drop_workspace_bytes": (SZ, [I64, I64, I]),
    "han_project_bwd_drop": (c_int, [P, I64, I64, I64, P, I, I, I, P, I64, P, SZ, P, FL, I, I64, P]),
    "han_reduce_blocks": (c_int, []),
    "han_attn_bwd_prep": (c_int, [P, I64, P, I64, P, P, I64, I, I, I, P, P, I64, P, P, P, P]),
    "han_attn_bwd_finish": (c_int, [P, I64, I, I, P, P, P, P, P, P, P, FL, I, I64, P]),
    "han_reduce_partials": (c_int, [P, I, I64, P, P]),
    "han_semantic_shape_supported": (c_int, [I, I]),
    "han_semantic_fwd": (c_int, [P, I64, I, I, I, P, P, P, I, P, P, P, P, P]),
    "han_semantic_combine": (c_int, [P, I64, I, I, P, P, P, P]),
    "han_semantic_bwd_workspace_bytes": (SZ, [I, I, I]),
    "han_semantic_bwd": (c_int, [P, P, P, P, I64, I, I, I, P, P, I, P, P, P, P, P, P, SZ, P, I64, P]),
    "han_semantic_tc_workspace_bytes": (SZ, []),
    "han_semantic_bwd_tc_workspace_bytes": (SZ, []),
    "han_semantic_bwd_tc": (c_int, [P, P, P, I64, I, I, I, P, P, P, I, P, P, P, P, P, P, SZ, P, I64, P]),
    "han_semantic_fwd_tc": (c_int, [P, I64, I, I, I, P, P, P, I, P, P, P, P, P, SZ, I, P]),
    "han_dense_blocks": (c_int, []),
    "han_dense_fwd": (c_int, [P, I64, I, I64, P, I, P, P, P]),
    "han_dense_bwd": (c_int, [P, I64, I, I64, P, I, P, P, P, P, P]),
    "han_masked_ce": (c_int, [P, P, P, P, I64, I, P, P, P]),
    "han_adam_l2_step": (c_int, [P, P, P, P, I64, P, FL, FL, FL, FL, FL, P]),
    "han_project_dx": (c_int, [P, I64, I, I, P, I64, I64, P, I64, I, P, FL, I, I64, P]),
}

_lib = None


def exported_symbols():
    return sorted(_SIGNATURES)


def load() -> ctypes.CDLL:
    """Loads the library (does not need a GPU; no CUDA call happens at load time)."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise ImportError(
            f"{_LIB_PATH} is missing: build it with `python -m han_b200.build` "
            "(han_b200 has no CPU or PyTorch fallback)")
    lib = ctypes.CDLL(str(_LIB_PATH))
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class HanError(RuntimeError):
    pass


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return c_void_p(t.data_ptr())


def stream_ptr():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


# kernels launched by one call of each entry point (for bench.py's `gpu_launches` count)
KERNELS_PER_CALL = {
    "han_dense_row_counts": 1, "han_scan_counts": 3, "han_dense_fill_indices": 1, "han_csr_transpose": 7,
    "han_csr_sort_rows": 2, "han_project_fwd": None, "han_project_bwd": 2,
    "han_attn_coefs": 1, "han_attn_bwd_prep": 1,
    "han_csr_chunk_rows": 1, "han_attn_fwd_chunked": 1, "han_attn_bwd_src_chunked": 1,
    "han_attn_bwd_finish": 1, "han_reduce_partials": 1, "han_semantic_fwd": 1, "han_semantic_combine": 1,
    "han_semantic_bwd": 2, "han_adam_l2_step": 1, "han_dense_fwd": 1, "han_dense_bwd": 1, "han_masked_ce": 1, "han_project_dx": 1,
    "han_attn_fwd_chunked_split": 2, "han_attn_bwd_src_chunked_split": 2, "han_semantic_fwd_tc": 2, "han_semantic_bwd_tc": 3,
}


class CallRecorder:
    """Optional per-entry-point CUDA-event timing and launch counting (used by bench.py only).
    Events are recorded on torch's current stream, the stream every kernel is launched on."""

    def __init__(self, time_events: bool = True):
        self.time_events = time_events
        self.events = {}      # name -> list of (start, end)
        self.launches = 0

    def summary(self):
        """name -> (calls, total_ms); call after a device synchronize."""
        return {k: (len(v), sum(s.elapsed_time(e) for s, e in v)) for k, v in self.events.items()}


_recorder = None

# Optional timeline (debugging aid, HAN_TRACE=1 in bench.py): (label, event) pairs recorded on whatever
# stream is current; elapsed times against the first event give a per-step timeline across streams.
TRACE = None


def trace_mark(label: str):
    if TRACE is not None:
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        TRACE.append((label, ev))


def set_recorder(rec):
    global _recorder
    _recorder = rec


def call(name: str, *args, kernels: int = None):
    """Calls an int-returning entry point; non-zero -> HanError(han_last_error())."""
    lib = load()
    rec = _recorder
    if rec is not None:
        k = KERNELS_PER_CALL.get(name) if kernels is None else kernels
        rec.launches += k if k is not None else 1
        if rec.time_events:
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
    if TRACE is not None:
        trace_mark(name + " >")
    rc = getattr(lib, name)(*args)
    if TRACE is not None:
        trace_mark(name + " <")
    if rec is not None and rec.time_events:
        e.record()
        rec.events.setdefault(name, []).append((s, e))
    if rc != 0:
        msg = lib.han_last_error()
        raise HanError(f"{name} failed (rc={rc}): {msg.decode() if msg else ''}")


def query(name: str, *args):
    """Calls a value-returning entry point (sizes, strides, flags)."""
    return getattr(load(), name)(*args)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise HanError("han_b200 ops run on CUDA tensors only (no CPU fallback)")
