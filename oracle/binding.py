"""ctypes access to the CPU checker libraries.  TEST INFRASTRUCTURE ONLY.

  libjpd_oracle.so       -- plain-C restatement of the reference algorithm (jpd_oracle.c)
  _ref/libjpd_ref.so     -- the UNMODIFIED reference header behind ref_shim.cpp (optional;
                            built in the container from /root/reference, travels prebuilt)

Allowed importers: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline and
--impl reference).  The product package jacobi_pd_b200 must never import this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libjpd_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libjpd_ref.so")
_cache = {}


def build(ref: bool = True) -> None:
    """(Re)build the checker; `ref` also rebuilds oracle/_ref when /root/reference exists."""
    targets = ["oracle"] + (["ref"] if ref else [])
    res = subprocess.run(["make", "-C", HERE] + targets, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + res.stdout + res.stderr)


def _load(path):
    if path not in _cache:
        _cache[path] = ctypes.CDLL(path)
    return _cache[path]


def oracle_lib():
    if not os.path.exists(ORACLE_SO):
        build(ref=False)
    return _load(ORACLE_SO)


def ref_available() -> bool:
    return os.path.exists(REF_SO)


def ref_lib():
    if not ref_available():
        raise FileNotFoundError(REF_SO + " (built only where /root/reference exists)")
    return _load(REF_SO)


def _sfx(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return "f64"
    if dtype == np.float32:
        return "f32"
    raise TypeError(dtype)


def _run(fn, mat, sort_criteria, calc_evecs, max_num_sweeps, n_threads):
    mat = np.ascontiguousarray(mat)
    batch, n, n2 = mat.shape
    assert n == n2
    evals = np.empty((batch, n), mat.dtype)
    evecs = np.empty((batch, n, n), mat.dtype) if calc_evecs else None
    iters = np.empty(batch, np.int32)
    fn.restype = ctypes.c_int
    used = fn(ctypes.c_void_p(mat.ctypes.data), ctypes.c_void_p(evals.ctypes.data),
              ctypes.c_void_p(evecs.ctypes.data) if calc_evecs else None,
              ctypes.c_void_p(iters.ctypes.data), ctypes.c_int(n), ctypes.c_longlong(batch),
              ctypes.c_int(sort_criteria), ctypes.c_int(1 if calc_evecs else 0),
              ctypes.c_int(max_num_sweeps), ctypes.c_int(n_threads))
    return evals, evecs, iters, used


def oracle_diagonalize(mat, sort_criteria=1, calc_evecs=True, max_num_sweeps=50, n_threads=0):
    """AoS batch (B,n,n) -> (evals (B,n), evecs (B,n,n) rows=eigenvectors, n_iters (B,), threads)."""
    fn = getattr(oracle_lib(), "jpdo_diagonalize_batched_" + _sfx(mat.dtype))
    return _run(fn, mat, sort_criteria, calc_evecs, max_num_sweeps, n_threads)


def ref_diagonalize(mat, sort_criteria=1, calc_evecs=True, max_num_sweeps=50, n_threads=0):
    """Same through the unmodified reference header (oracle/_ref)."""
    fn = getattr(ref_lib(), "jpdref_diagonalize_batched_" + _sfx(mat.dtype))
    return _run(fn, mat, sort_criteria, calc_evecs, max_num_sweeps, n_threads)


def best_cpu_diagonalize(mat, **kw):
    """The reference itself when present, else the pinned restatement; returns (..., kind)."""
    if ref_available():
        return ref_diagonalize(mat, **kw) + ("reference",)
    return oracle_diagonalize(mat, **kw) + ("port",)


def host_threads() -> int:
    fn = oracle_lib().jpdo_host_threads
    fn.restype = ctypes.c_int
    return fn()


def fill_synthetic(n, batch, layout=0, kind=0, seed=1234, b0=0, dtype=np.float64):
    """Host twin of jpd_fill_synthetic_* -- identical bits to the device generator."""
    shape = (batch, n, n) if layout == 0 else (n, n, batch)
    out = np.empty(shape, dtype)
    fn = getattr(oracle_lib(), "jpdo_fill_synthetic_" + _sfx(dtype))
    fn.restype = ctypes.c_int
    rc = fn(ctypes.c_void_p(out.ctypes.data), ctypes.c_int(n), ctypes.c_longlong(batch),
            ctypes.c_int(layout), ctypes.c_int(kind), ctypes.c_ulonglong(seed), ctypes.c_longlong(b0))
    if rc != 0:
        raise ValueError("jpdo_fill_synthetic: bad arguments")
    return out
