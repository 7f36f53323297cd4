"""CPU-only checks of the drop-in boundary: libjpd.so loads, exports every symbol that
include/jpd.h (product ABI) and include/jpd_bench.h (bench/test scaffolding) declare, rejects bad arguments before touching the GPU, and the pure host
helpers (slicing, byte model, tier choice) behave.  No compute call is made here."""
import ctypes
import os
import re

import pytest

import jacobi_pd_b200 as jpd
from jacobi_pd_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(headers=("jpd.h", "jpd_bench.h")):
    text = ""
    for h in headers:
        with open(os.path.join(ROOT, "include", h)) as f:
            text += f.read()
    return sorted(set(re.findall(r"JPD_API\s+[\w\s\*]+?\b(jpd_\w+)\s*\(", text)))


def test_library_is_built_in_tree():
    assert os.path.exists(jpd.LIB_PATH) and jpd.LIB_PATH.startswith(ROOT)


def test_every_declared_symbol_is_exported_and_bound():
    syms = declared_symbols()
    assert len(syms) >= 15 and "jpd_diagonalize_batched_f64" in syms and "jpd_diagonalize_batched_f32" in syms
    handle = ctypes.CDLL(jpd.LIB_PATH)
    for s in syms:
        assert getattr(handle, s) is not None
    assert sorted(_lib.SIGNATURES) == syms      # the Python binding covers exactly the headers
    # the product header carries no bench scaffolding
    assert not [s for s in declared_symbols(("jpd.h",)) if "synthetic" in s or "probe" in s]


def test_header_cites_the_reference_interface():
    with open(os.path.join(ROOT, "include", "jpd.h")) as f:
        text = f.read()
    assert "jacobi_pd.hpp:76-81" in text and "jacobi_pd.hpp:159-220" in text


def test_argument_validation_needs_no_gpu():
    l = jpd.lib()
    f = l.jpd_diagonalize_batched_f64
    assert f(None, None, None, None, 0, 10, 0, 1, 1, 50, None) == -1          # n < 1
    assert f(None, None, None, None, 4, -1, 0, 1, 1, 50, None) == -1          # negative batch
    assert f(None, None, None, None, 4, 10, 7, 1, 1, 50, None) == -1          # unknown layout
    assert f(None, None, None, None, 4, 10, 0, 1, 1, 50, None) == -1          # null pointers
    assert f(None, None, None, None, 5000, 10, 0, 1, 1, 50, None) == -1       # n > JPD_MAX_N
    assert f(None, None, None, None, 4, 0, 0, 1, 1, 50, None) == 0            # empty batch is a no-op
    assert l.jpd_diagonalize_batched_host_f64(None, None, None, None, 4, 0, 0, 1, 1, 50, 0) == 0
    assert l.jpd_diagonalize_batched_host_f32(None, None, None, None, 4, 5, 0, 1, 1, 50, 0) == -1
    assert l.jpd_fill_synthetic_f64(None, 4, 5, 0, 0, 1, 0, None) == -1
    assert l.jpd_error_string(-1).decode() == "invalid argument"
    assert l.jpd_version() >= 100


def test_slices_partition_the_batch():
    for batch in (0, 1, 7, 1000, 16777216, 16777217):
        for parts in (1, 2, 3, 4, 8):
            got, end = 0, 0
            for p in range(parts):
                first, count = jpd.slice_of(batch, parts, p)
                assert first == end or count == 0
                end = max(end, first + count)
                got += count
            assert got == batch and end == batch


def test_byte_model_and_tiers():
    assert jpd.algorithmic_bytes(3, 8) == 148 and jpd.algorithmic_bytes(4, 8) == 244
    assert jpd.algorithmic_bytes(32, 8) == 12676 and jpd.algorithmic_bytes(3, 4) == 76
    assert jpd.algorithmic_bytes(4, 8, False) == 244 - 128
    l = jpd.lib()
    assert [l.jpd_kernel_tier(n, 8, 1) for n in (1, 3, 4, 6)] == [1, 1, 1, 1]
    assert l.jpd_kernel_tier(32, 8, 1) == 2 and l.jpd_kernel_tier(64, 8, 1) == 3 and l.jpd_kernel_tier(128, 8, 1) == 3 and l.jpd_kernel_tier(129, 8, 1) == 4 and l.jpd_kernel_tier(256, 8, 1) == 4 and l.jpd_kernel_tier(257, 8, 1) == 5


def test_no_silent_cpu_fallback():
    """Without a CUDA tensor the Python front end refuses instead of computing on the host."""
    import torch
    with pytest.raises(jpd.JpdError):
        jpd.diagonalize_batched(torch.zeros(2, 3, 3, dtype=torch.float64))
    import inspect
    src = inspect.getsource(jpd.api) + inspect.getsource(_lib)
    assert "oracle" not in src.replace("never imported", "")     # product never reaches into oracle/


def test_python_binding_validates_host_buffers_before_the_call():
    """A wrong dtype / size / stride would be a silent out-of-bounds write behind raw pointers."""
    import numpy as np
    n, B = 3, 4
    mat = np.zeros((B, n, n)); ev = np.zeros((B, n)); vec = np.zeros((B, n, n)); it = np.zeros(B, np.int32)
    with pytest.raises(ValueError, match="dtype"):
        jpd.diagonalize_batched_host(mat.astype(np.float32), ev, vec, it, n, B)
    with pytest.raises(ValueError, match="dtype"):
        jpd.diagonalize_batched_host(mat, ev, vec, it.astype(np.int64), n, B)
    with pytest.raises(ValueError, match="elements"):
        jpd.diagonalize_batched_host(mat, ev[:2], vec, it, n, B)
    with pytest.raises(ValueError, match="contiguous"):
        jpd.diagonalize_batched_host(mat, ev, np.zeros((B, n, 2 * n))[:, :, ::2], it, n, B)
    with pytest.raises(ValueError, match="required"):
        jpd.diagonalize_batched_host(mat, ev, None, it, n, B)
    with pytest.raises(ValueError, match="elements"):
        jpd.quaternion_from_correlation_host(np.zeros((B, 8)), np.zeros(B), np.zeros((B, 4)), it, B)
    l = jpd.lib()
    assert l.jpd_quaternion_from_correlation_host_f64(None, None, None, None, 5, 0, 0, 50, 0) == -1
    assert l.jpd_principal_axes_host_f32(None, None, None, None, 0, 0, 50, 0) == 0
    assert l.jpd_diagonalize_batched_host_packed_f64(None, None, None, None, 4, 5, 0, 1, 1, 50, 0) == -1
    assert l.jpd_host_register(None, 16) == -1


def test_python_jacobi_class_fills_lists_in_place():
    """Jacobi.Diagonalize writes into whatever [i] / [i][j] container the caller passes; the
    marshalling is checked without a GPU by stubbing the host entry."""
    import numpy as np
    from jacobi_pd_b200 import api
    calls = {}

    def fake(mat, ev, vec, it, n, batch, *a, **k):
        calls["n"] = n
        ev[...] = np.arange(n, 0, -1); vec[...] = np.eye(n); it[0] = 7
    orig, api.diagonalize_batched_host = api.diagonalize_batched_host, fake
    try:
        evals, evecs = [0.0] * 3, [[0.0] * 3 for _ in range(3)]
        assert jpd.Jacobi(3).Diagonalize([[2, 1, 1], [1, 2, -1], [1, -1, 2]], evals, evecs) == 7
        assert evals == [3.0, 2.0, 1.0] and evecs[1] == [0.0, 1.0, 0.0] and calls["n"] == 3
    finally:
        api.diagonalize_batched_host = orig
