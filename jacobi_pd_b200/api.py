"""Host-side mirror of the reference interface for the Diagonalize path.

Reference: ``jacobi_pd::Jacobi<Scalar,Vector,Matrix,ConstMatrix>`` in
/root/reference/include/jacobi_pd.hpp:25-146 (``Diagonalize`` :76-81, ``SortCriteria``
:56-62, ``SetSize`` :42).  The real drop-in for C++ callers is
``include/jacobi_pd_cuda.hpp``; this module is the same surface for Python callers,
tests and ``bench.py`` -- thin marshalling over the C ABI, PyTorch only as the owner of
device memory and streams.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib

# values as the reference enum (jacobi_pd.hpp:56-62)
DO_NOT_SORT = 0
SORT_DECREASING_EVALS = 1
SORT_INCREASING_EVALS = 2
SORT_DECREASING_ABS_EVALS = 3
SORT_INCREASING_ABS_EVALS = 4

LAYOUT_AOS = 0
LAYOUT_SOA = 1
LAYOUT_PACKED_AOS = 2   # adapter-only layouts, see include/jpd.h
LAYOUT_PACKED_SOA = 3
LAYOUT_POINTERS = 4


def _suffix(dtype) -> str:
    import torch
    if dtype in (torch.float64, np.float64, np.dtype("float64")):
        return "f64"
    if dtype in (torch.float32, np.float32, np.dtype("float32")):
        return "f32"
    raise TypeError(f"unsupported dtype {dtype}: the path computes in fp64 or fp32")


def out_shapes(n: int, batch: int, layout: int):
    if layout == LAYOUT_AOS:
        return (batch, n), (batch, n, n)
    return (n, batch), (n, n, batch)


def diagonalize_batched(mat, n: int | None = None, layout: int = LAYOUT_AOS,
                        sort_criteria: int = SORT_DECREASING_EVALS, calc_evecs: bool = True,
                        max_num_sweeps: int = 50, out=None, stream=None):
    """Diagonalise a device-resident batch (torch CUDA tensor) through the C ABI.

    mat: AoS ``(batch, n, n)`` or SoA-interleaved ``(n, n, batch)`` contiguous CUDA tensor.
    Returns ``(evals, evecs, n_iters)`` device tensors (evecs is None when not requested);
    eigenvector k is row k (``evecs[b, k, :]`` / ``evecs[k, :, b]``).  Asynchronous on the
    current torch stream.
    """
    import torch
    if not mat.is_cuda:
        raise _lib.JpdError("diagonalize_batched needs a CUDA tensor (no CPU fallback); "
                            "use diagonalize_batched_host for host memory")
    if not mat.is_contiguous():
        raise ValueError("mat must be contiguous")
    if layout == LAYOUT_AOS:
        batch, n_, n2 = mat.shape
    else:
        n_, n2, batch = mat.shape
    if n_ != n2 or (n is not None and n != n_):
        raise ValueError("mat must hold square n x n matrices")
    n = n_
    sfx = _suffix(mat.dtype)
    es, vs = out_shapes(n, batch, layout)
    if out is None:
        evals = torch.empty(es, dtype=mat.dtype, device=mat.device)
        evecs = torch.empty(vs, dtype=mat.dtype, device=mat.device) if calc_evecs else None
        iters = torch.empty((batch,), dtype=torch.int32, device=mat.device)
    else:
        evals, evecs, iters = out
    st = torch.cuda.current_stream(mat.device).cuda_stream if stream is None else stream
    with torch.cuda.device(mat.device):
        fn = getattr(_lib.lib(), f"jpd_diagonalize_batched_{sfx}")
        rc = fn(mat.data_ptr(), evals.data_ptr(), evecs.data_ptr() if evecs is not None else None,
                iters.data_ptr() if iters is not None else None, n, batch, layout, sort_criteria,
                1 if calc_evecs else 0, max_num_sweeps, st)
    _lib.check(rc, f"jpd_diagonalize_batched_{sfx}")
    return evals, evecs, iters


def quaternion_from_correlation(corr, layout: int = LAYOUT_AOS, all_eigenpairs: bool = False,
                                max_num_sweeps: int = 50, stream=None, out=None):
    """colvars optimal-rotation path fused in one kernel: 3x3 correlation matrices (CUDA
    tensor, AoS ``(batch, 9)``/``(batch, 3, 3)`` or SoA ``(9, batch)``) -> ``(lambda, quat,
    n_iters)``; with ``all_eigenpairs`` the four eigenvalues (decreasing) and the 4x4
    eigenvector rows of the overlap matrix, laid out as ``diagonalize_batched`` would."""
    import torch
    if not corr.is_cuda or not corr.is_contiguous():
        raise _lib.JpdError("quaternion_from_correlation needs a contiguous CUDA tensor (no CPU fallback)")
    batch = corr.shape[0] if layout == LAYOUT_AOS else corr.shape[-1]
    if corr.numel() != 9 * batch:
        raise ValueError("corr must hold 9 scalars per matrix")
    sfx = _suffix(corr.dtype)
    if all_eigenpairs:
        es, vs = out_shapes(4, batch, layout)
    else:
        es, vs = ((batch,), (batch, 4)) if layout == LAYOUT_AOS else ((batch,), (4, batch))
    if out is None:
        lam = torch.empty(es, dtype=corr.dtype, device=corr.device)
        quat = torch.empty(vs, dtype=corr.dtype, device=corr.device)
        iters = torch.empty((batch,), dtype=torch.int32, device=corr.device)
    else:                                   # reuse the caller's (lambda, quat, n_iters) buffers
        lam, quat, iters = out
    st = torch.cuda.current_stream(corr.device).cuda_stream if stream is None else stream
    with torch.cuda.device(corr.device):
        rc = getattr(_lib.lib(), f"jpd_quaternion_from_correlation_{sfx}")(
            corr.data_ptr(), lam.data_ptr(), quat.data_ptr(), iters.data_ptr(), batch, layout,
            1 if all_eigenpairs else 0, max_num_sweeps, st)
    _lib.check(rc, "jpd_quaternion_from_correlation")
    return lam, quat, iters


def principal_axes(inertia6, layout: int = LAYOUT_AOS, max_num_sweeps: int = 50, stream=None, out=None):
    """LAMMPS rigid-body path fused in one kernel: (Ixx, Iyy, Izz, Iyz, Ixz, Ixy) per body
    (CUDA tensor, AoS ``(batch, 6)`` or SoA ``(6, batch)``) -> ``(moments, axes, n_iters)``:
    moments decreasing, ``axes[b, i, k]`` = component i of principal axis k, right-handed."""
    import torch
    if not inertia6.is_cuda or not inertia6.is_contiguous():
        raise _lib.JpdError("principal_axes needs a contiguous CUDA tensor (no CPU fallback)")
    batch = inertia6.shape[0] if layout == LAYOUT_AOS else inertia6.shape[-1]
    if inertia6.numel() != 6 * batch:
        raise ValueError("inertia6 must hold 6 scalars per body")
    sfx = _suffix(inertia6.dtype)
    es, vs = out_shapes(3, batch, layout)
    if out is None:
        mom = torch.empty(es, dtype=inertia6.dtype, device=inertia6.device)
        axes = torch.empty(vs, dtype=inertia6.dtype, device=inertia6.device)
        iters = torch.empty((batch,), dtype=torch.int32, device=inertia6.device)
    else:
        mom, axes, iters = out
    st = torch.cuda.current_stream(inertia6.device).cuda_stream if stream is None else stream
    with torch.cuda.device(inertia6.device):
        rc = getattr(_lib.lib(), f"jpd_principal_axes_{sfx}")(
            inertia6.data_ptr(), mom.data_ptr(), axes.data_ptr(), iters.data_ptr(), batch, layout,
            max_num_sweeps, st)
    _lib.check(rc, "jpd_principal_axes")
    return mom, axes, iters


def layout_shape(n: int, batch: int, layout: int):
    t = n * (n + 1) // 2
    return {LAYOUT_AOS: (batch, n, n), LAYOUT_SOA: (n, n, batch), LAYOUT_PACKED_AOS: (batch, t),
            LAYOUT_PACKED_SOA: (t, batch)}[layout]


def convert_layout(src, src_layout: int, dst_layout: int, n: int, batch: int, dtype=None, stream=None):
    """Layout adapter: CUDA tensor in `src_layout` (for LAYOUT_POINTERS an
    int64 tensor of device addresses, `dtype` giving the scalar type) -> new tensor in
    `dst_layout`; one tiled-transpose kernel."""
    import torch
    if not src.is_cuda or not src.is_contiguous():
        raise _lib.JpdError("convert_layout needs a contiguous CUDA tensor (no CPU fallback)")
    dt = dtype if src_layout == LAYOUT_POINTERS else src.dtype
    sfx = _suffix(dt)
    dst = torch.empty(layout_shape(n, batch, dst_layout), dtype=dt, device=src.device)
    st = torch.cuda.current_stream(src.device).cuda_stream if stream is None else stream
    with torch.cuda.device(src.device):
        rc = getattr(_lib.lib(), f"jpd_convert_layout_{sfx}")(src.data_ptr(), src_layout, dst.data_ptr(),
                                                              dst_layout, n, batch, st)
    _lib.check(rc, "jpd_convert_layout")
    return dst


def _host_ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()  # torch CPU tensor (possibly pinned)


def _host_check(name, a, dtype, count, optional=False):
    """The C ABI trusts raw host pointers: a wrong dtype / size / stride would be a silent
    out-of-bounds write, so everything is validated here (ValueError, nothing is written)."""
    if a is None:
        if optional:
            return
        raise ValueError(f"{name} is required")
    if isinstance(a, np.ndarray):
        ok_dtype, contiguous, size = a.dtype == np.dtype(dtype), a.flags["C_CONTIGUOUS"], a.size
    else:  # torch CPU tensor
        import torch
        if a.is_cuda:
            raise ValueError(f"{name} must be a host array (use diagonalize_batched for CUDA tensors)")
        want = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32,
                np.dtype(np.int32): torch.int32}[np.dtype(dtype)]
        ok_dtype, contiguous, size = a.dtype == want, a.is_contiguous(), a.numel()
    if not ok_dtype:
        raise ValueError(f"{name} must have dtype {np.dtype(dtype).name}")
    if not contiguous:
        raise ValueError(f"{name} must be C-contiguous")
    if size != count:
        raise ValueError(f"{name} must hold {count} elements, has {size}")


def _np_dtype(a):
    return a.dtype if isinstance(a, np.ndarray) else {"torch.float64": np.dtype(np.float64),
                                                       "torch.float32": np.dtype(np.float32)}[str(a.dtype)]


def diagonalize_batched_host(mat, evals, evecs, n_iters, n: int, batch: int, layout: int = LAYOUT_AOS,
                             sort_criteria: int = SORT_DECREASING_EVALS, calc_evecs: bool = True,
                             max_num_sweeps: int = 50, device: int = 0, n_devices: int = 1, devices=None,
                             packed: bool = False):
    """Host buffers in, host buffers out (numpy arrays or CPU torch tensors, written in place).
    ``packed``: `mat` holds only the upper triangle, n(n+1)/2 scalars per matrix
    (jpd_diagonalize_batched_host_packed_*).  Several GPUs: ``devices=[...]`` (or
    ``n_devices`` for devices 0..n_devices-1); `device` is then not used."""
    dt = _np_dtype(mat)
    sfx = _suffix(dt)
    per_in = n * (n + 1) // 2 if packed else n * n
    _host_check("mat", mat, dt, per_in * batch)
    _host_check("evals", evals, dt, n * batch)
    _host_check("evecs", evecs, dt, n * n * batch, optional=not calc_evecs)
    _host_check("n_iters", n_iters, np.int32, batch, optional=True)
    l = _lib.lib()
    vec_ptr = _host_ptr(evecs) if calc_evecs else None
    if devices is not None:
        n_devices = len(devices)
    if packed:
        if n_devices > 1:
            raise ValueError("packed input is a single-device entry")
        fn = getattr(l, f"jpd_diagonalize_batched_host_packed_{sfx}")
        rc = fn(_host_ptr(mat), _host_ptr(evals), vec_ptr, _host_ptr(n_iters), n, batch, layout,
                sort_criteria, 1 if calc_evecs else 0, max_num_sweeps, device)
    elif n_devices > 1 or devices is not None:
        fn = getattr(l, f"jpd_diagonalize_batched_host_multi_{sfx}")
        dev_arr = (ctypes.c_int * n_devices)(*devices) if devices is not None else None
        rc = fn(_host_ptr(mat), _host_ptr(evals), vec_ptr, _host_ptr(n_iters), n, batch, layout,
                sort_criteria, 1 if calc_evecs else 0, max_num_sweeps, n_devices, dev_arr)
    else:
        fn = getattr(l, f"jpd_diagonalize_batched_host_{sfx}")
        rc = fn(_host_ptr(mat), _host_ptr(evals), vec_ptr, _host_ptr(n_iters), n, batch, layout,
                sort_criteria, 1 if calc_evecs else 0, max_num_sweeps, device)
    _lib.check(rc, "jpd_diagonalize_batched_host")


def quaternion_from_correlation_host(corr, lam, quat, n_iters, batch: int, layout: int = LAYOUT_AOS,
                                     all_eigenpairs: bool = False, max_num_sweeps: int = 50, device: int = 0):
    """colvars path with HOST buffers (written in place): 9 scalars up, 1 + 4 (or 4 + 16) down."""
    dt = _np_dtype(corr)
    _host_check("corr", corr, dt, 9 * batch)
    _host_check("lam", lam, dt, (4 if all_eigenpairs else 1) * batch)
    _host_check("quat", quat, dt, (16 if all_eigenpairs else 4) * batch)
    _host_check("n_iters", n_iters, np.int32, batch, optional=True)
    rc = getattr(_lib.lib(), f"jpd_quaternion_from_correlation_host_{_suffix(dt)}")(
        _host_ptr(corr), _host_ptr(lam), _host_ptr(quat), _host_ptr(n_iters), batch, layout,
        1 if all_eigenpairs else 0, max_num_sweeps, device)
    _lib.check(rc, "jpd_quaternion_from_correlation_host")


def principal_axes_host(inertia6, moments, axes, n_iters, batch: int, layout: int = LAYOUT_AOS,
                        max_num_sweeps: int = 50, device: int = 0):
    """LAMMPS path with HOST buffers (written in place): 6 scalars up, 3 + 9 down."""
    dt = _np_dtype(inertia6)
    _host_check("inertia6", inertia6, dt, 6 * batch)
    _host_check("moments", moments, dt, 3 * batch)
    _host_check("axes", axes, dt, 9 * batch)
    _host_check("n_iters", n_iters, np.int32, batch, optional=True)
    rc = getattr(_lib.lib(), f"jpd_principal_axes_host_{_suffix(dt)}")(
        _host_ptr(inertia6), _host_ptr(moments), _host_ptr(axes), _host_ptr(n_iters), batch, layout,
        max_num_sweeps, device)
    _lib.check(rc, "jpd_principal_axes_host")


def host_register(a) -> None:
    """Page-lock a caller-owned numpy array so that the host entries overlap copies and kernels."""
    _lib.check(_lib.lib().jpd_host_register(_host_ptr(a), a.nbytes), "jpd_host_register")


def host_unregister(a) -> None:
    _lib.check(_lib.lib().jpd_host_unregister(_host_ptr(a)), "jpd_host_unregister")


def fill_synthetic(mat, n: int, batch: int, layout: int, kind: int, seed: int, b0: int = 0):
    """Fill a CUDA tensor with the counter-based synthetic batch."""
    import torch
    sfx = _suffix(mat.dtype)
    st = torch.cuda.current_stream(mat.device).cuda_stream
    with torch.cuda.device(mat.device):
        rc = getattr(_lib.lib(), f"jpd_fill_synthetic_{sfx}")(mat.data_ptr(), n, batch, layout, kind, seed, b0, st)
    _lib.check(rc, "jpd_fill_synthetic")
    return mat


def slice_of(batch: int, n_parts: int, part: int):
    """(first, count) of part `part` of `n_parts` -- the multi-GPU partition."""
    first, count = ctypes.c_longlong(), ctypes.c_longlong()
    _lib.lib().jpd_slice(batch, n_parts, part, ctypes.byref(first), ctypes.byref(count))
    return first.value, count.value


def algorithmic_bytes(n: int, elem_size: int, calc_evecs: bool = True) -> int:
    return _lib.lib().jpd_algorithmic_bytes(n, elem_size, 1 if calc_evecs else 0)


class Jacobi:
    """Python twin of ``jacobi_pd::Jacobi`` (jacobi_pd.hpp:25-146) for ONE matrix at a time.

    ``Diagonalize(mat, evals, evecs, sort_criteria, calc_evecs, max_num_sweeps)`` fills
    the caller's ``evals`` (n) and ``evecs`` (n x n, eigenvector k in row k) numpy arrays in
    place and returns the number of iterations (>0) or 0 on non-convergence, exactly the
    reference contract (:64-81).  The work is done on the GPU through the host entry
    point of the C ABI (batch = 1); a CUDA device is required.
    """
    DO_NOT_SORT = DO_NOT_SORT
    SORT_DECREASING_EVALS = SORT_DECREASING_EVALS
    SORT_INCREASING_EVALS = SORT_INCREASING_EVALS
    SORT_DECREASING_ABS_EVALS = SORT_DECREASING_ABS_EVALS
    SORT_INCREASING_ABS_EVALS = SORT_INCREASING_ABS_EVALS

    def __init__(self, n: int = 0, dtype=np.float64, device: int = 0):
        self.dtype = np.dtype(dtype)
        self.device = device
        self.SetSize(n)

    def SetSize(self, n: int) -> None:
        self.n = int(n)

    def Diagonalize(self, mat, evals, evecs, sort_criteria: int = SORT_DECREASING_EVALS,
                    calc_evecs: bool = True, max_num_sweeps: int = 50) -> int:
        n = self.n
        m = np.ascontiguousarray(np.asarray(mat, dtype=self.dtype).reshape(n, n))
        ev = np.empty(n, self.dtype)
        vec = np.empty((n, n), self.dtype) if calc_evecs else None
        it = np.zeros(1, np.int32)
        diagonalize_batched_host(m, ev, vec, it, n, 1, LAYOUT_AOS, sort_criteria, calc_evecs,
                                 max_num_sweeps, self.device)
        # the caller's containers are filled in place: numpy arrays, or anything [i] / [i][j]
        # indexable as in the reference (lists of lists included)
        if isinstance(evals, np.ndarray):
            evals[...] = ev.reshape(evals.shape)
        else:
            for i in range(n):
                evals[i] = float(ev[i])
        if calc_evecs:
            if isinstance(evecs, np.ndarray):
                evecs[...] = vec.reshape(evecs.shape)
            else:
                for i in range(n):
                    for j in range(n):
                        evecs[i][j] = float(vec[i, j])
        return int(it[0])
