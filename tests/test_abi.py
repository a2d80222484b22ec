"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every
symbol include/han_b200.h declares (and nothing the ctypes table does not know), and the pure
host queries answer.  No compute entry point is called here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "han_b200.h")


@pytest.fixture(scope="module")
def lib():
    from han_b200 import build, _lib
    build.build(verbose=False)          # no-op when up to date; nvcc cross-compiles without a GPU
    return _lib.load()


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(han_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_what_ctypes_binds(lib):
    from han_b200 import _lib
    assert header_symbols() == _lib.exported_symbols()


def test_library_exports_every_header_symbol(lib):
    raw = ctypes.CDLL(os.path.join(ROOT, "han_b200", "libhan_sm100.so"))
    for name in header_symbols():
        assert hasattr(raw, name), f"{name} declared in han_b200.h but not exported"


def test_host_queries(lib):
    assert lib.han_version() >= 100
    assert lib.han_attn_shape_supported(8, 8) == 1
    assert lib.han_attn_shape_supported(3, 5) == 0
    assert lib.han_table_stride(8, 8) == 64 and lib.han_record_stride(8, 8) == 88
    assert lib.han_table_stride(1, 8) == 8 and lib.han_record_stride(1, 8) == 12
    assert lib.han_csr_num_chunks(0) == 1 and lib.han_csr_num_chunks(5000) == 5000 // 128 + 1
    assert lib.han_csr_chunk_edges(100_000_000) == 2048 and lib.han_csr_chunk_edges(2_000_000) == 422
    assert lib.han_semantic_shape_supported(64, 128) == 1
    assert lib.han_reduce_blocks() > 0
    assert lib.han_scan_workspace_bytes(2_000_000) >= 8 * (2_000_000 // 2048)
    assert lib.han_project_bwd_workspace_bytes(2_000_000, 256, 4, 64) > 0
    assert lib.han_semantic_bwd_workspace_bytes(4, 64, 128) > 0


def test_invalid_arguments_return_negative_and_set_message(lib):
    # argument validation happens before any CUDA call, so this is safe without a GPU
    rc = lib.han_attn_fwd_chunked(None, None, None, 0, 0, None, None, None, None, None, 8, 8, 1, None, 64, None, None, None,
                                  None, 0, None, 0, 0, None, None, None, 1.0, 1.0, 0, 0, None)
    assert rc < 0
    assert b"han_attn_fwd_chunked" in lib.han_last_error()
    rc = lib.han_dense_row_counts(None, 0, 0, 4, 4, None, None, None)
    assert rc < 0


def test_product_has_no_cpu_fallback():
    import torch
    import han_b200 as hb
    x = torch.zeros(1, 4, 8)
    bias = torch.zeros(1, 4, 4)
    with pytest.raises(Exception):
        hb.layers.attn_head(x, 8, bias, hb.layers.elu)   # CPU tensors are refused, not emulated


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "han_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S).replace("#", "\n#").split("\n#")[0] or \
                not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn
