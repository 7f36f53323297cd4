"""-m gpu: the host-buffer entries of the C ABI (include/jpd.h) -- fused producers, packed
upper-triangle input, page-locking helper -- against the device entries (bit-identical) and
the oracle, through both the small-call path and the chunked three-stream pipeline."""
import os

import numpy as np
import pytest

from oracle import binding as ob
import parity
from gpu_util import run_gpu

pytestmark = pytest.mark.gpu


@pytest.fixture
def small_chunks(monkeypatch):
    monkeypatch.setenv("JPD_HOST_CHUNK_MB", "1")     # force many chunks at test sizes
    yield


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("batch", [7, 40000])          # small-call path / chunked pipeline
def test_quaternion_host_equals_device_entry(dtype, layout, batch, small_chunks):
    import torch
    import jacobi_pd_b200 as jpd
    rng = np.random.default_rng(21)
    R = rng.uniform(-1, 1, (batch, 9)).astype(dtype)
    src = np.ascontiguousarray(R if layout == 0 else R.T)
    for all_pairs in (False, True):
        lam_d, quat_d, it_d = jpd.quaternion_from_correlation(torch.from_numpy(src).cuda(), layout=layout,
                                                              all_eigenpairs=all_pairs)
        torch.cuda.synchronize()
        lam = np.empty(tuple(lam_d.shape), dtype); quat = np.empty(tuple(quat_d.shape), dtype)
        it = np.empty(batch, np.int32)
        jpd.quaternion_from_correlation_host(src, lam, quat, it, batch, layout, all_pairs)
        assert np.array_equal(lam, lam_d.cpu().numpy()) and np.array_equal(quat, quat_d.cpu().numpy())
        assert np.array_equal(it, it_d.cpu().numpy()) and (it > 0).all()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("batch", [5, 60000])
def test_principal_axes_host_equals_device_entry(dtype, layout, batch, small_chunks):
    import torch
    import jacobi_pd_b200 as jpd
    rng = np.random.default_rng(22)
    pts = rng.normal(size=(batch, 6, 3))
    r2 = (pts ** 2).sum(axis=2)
    I = np.einsum("bk,ij->bij", r2, np.eye(3)) - np.einsum("bki,bkj->bij", pts, pts)
    I6 = np.stack([I[:, 0, 0], I[:, 1, 1], I[:, 2, 2], I[:, 1, 2], I[:, 0, 2], I[:, 0, 1]], axis=1).astype(dtype)
    src = np.ascontiguousarray(I6 if layout == 0 else I6.T)
    mom_d, axes_d, it_d = jpd.principal_axes(torch.from_numpy(src).cuda(), layout=layout)
    torch.cuda.synchronize()
    mom = np.empty(tuple(mom_d.shape), dtype); axes = np.empty(tuple(axes_d.shape), dtype)
    it = np.empty(batch, np.int32)
    jpd.principal_axes_host(src, mom, axes, it, batch, layout)
    assert np.array_equal(mom, mom_d.cpu().numpy()) and np.array_equal(axes, axes_d.cpu().numpy())
    assert np.array_equal(it, it_d.cpu().numpy())
    # and against the oracle on the explicit tensors
    ref = ob.oracle_diagonalize(np.ascontiguousarray(I.astype(dtype)), 1)
    parity.check_evals(mom if layout == 0 else np.ascontiguousarray(mom.T), ref[0])


def pack_upper(A):
    n = A.shape[-1]
    iu = np.triu_indices(n)
    return np.ascontiguousarray(A[:, iu[0], iu[1]])


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("n,batch", [(3, 9), (4, 50000), (4, 11), (12, 3000), (40, 200)])
def test_packed_host_input_equals_full_input(n, batch, layout, dtype, small_chunks):
    import jacobi_pd_b200 as jpd
    A = ob.fill_synthetic(n, batch, seed=40 + n, dtype=dtype)
    A[:, np.tril_indices(n, -1)[0], np.tril_indices(n, -1)[1]] = 99.0    # lower triangle is never read (:167-171)
    full = run_gpu(A, 1, layout=layout, host_api=True)
    P = pack_upper(A)
    src = np.ascontiguousarray(P if layout == 0 else P.T)
    es, vs = jpd.api.out_shapes(n, batch, layout)
    ev = np.empty(es, dtype); vec = np.empty(vs, dtype); it = np.empty(batch, np.int32)
    jpd.diagonalize_batched_host(src, ev, vec, it, n, batch, layout, 1, True, 50, packed=True)
    if layout == 1:
        ev, vec = np.ascontiguousarray(ev.T), np.ascontiguousarray(vec.transpose(2, 0, 1))
    assert np.array_equal(ev, full[0]) and np.array_equal(vec, full[1]) and np.array_equal(it, full[2])
    parity.full_check(A, (ev, vec, it), ob.oracle_diagonalize(A, 1), 1)


def test_registered_pageable_buffers_and_reuse(small_chunks):
    import jacobi_pd_b200 as jpd
    n, batch = 4, 30000
    A = ob.fill_synthetic(n, batch, kind=1, seed=5)
    ev = np.empty((batch, n)); vec = np.empty((batch, n, n)); it = np.empty(batch, np.int32)
    for buf in (A, ev, vec):
        jpd.host_register(buf)
    try:
        for _ in range(2):     # slots and streams are reused by the second call
            ev[...] = 0; vec[...] = 0
            jpd.diagonalize_batched_host(A, ev, vec, it, n, batch)
            parity.full_check(A, (ev, vec, it), ob.oracle_diagonalize(A, 1), 1)
    finally:
        for buf in (A, ev, vec):
            jpd.host_unregister(buf)


def test_host_entry_restores_the_callers_device():
    import torch
    import jacobi_pd_b200 as jpd
    # with two GPUs: the caller sits on device 1 and computes on device 0; with one GPU the
    # same call must at least leave device 0 current (the DeviceGuard path is the same)
    here = 1 if torch.cuda.device_count() >= 2 else 0
    torch.cuda.set_device(here)
    A = ob.fill_synthetic(3, 100, seed=1)
    ev = np.empty((100, 3)); vec = np.empty((100, 3, 3)); it = np.empty(100, np.int32)
    jpd.diagonalize_batched_host(A, ev, vec, it, 3, 100, device=0)
    assert torch.cuda.current_device() == here
    torch.cuda.set_device(0)
    parity.full_check(A, (ev, vec, it), ob.oracle_diagonalize(A, 1), 1)
