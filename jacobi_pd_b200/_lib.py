"""ctypes binding of libjpd.so (C ABI in include/jpd.h).

The shared library is built in-tree by ``jacobi_pd_b200/csrc/Makefile`` (or
``__graft_entry__.build()``).  There is no fallback of any kind: if the library is
missing, or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libjpd.so")
_lib = None

c_int, c_ll, c_ull, c_vp = ctypes.c_int, ctypes.c_longlong, ctypes.c_ulonglong, ctypes.c_void_p

# name -> (restype, argtypes); must list every symbol declared in include/jpd.h and jpd_bench.h
SIGNATURES = {
    "jpd_diagonalize_batched_f64": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_ll, c_int, c_int, c_int, c_int, c_vp]),
    "jpd_diagonalize_batched_f32": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_ll, c_int, c_int, c_int, c_int, c_vp]),
    "jpd_diagonalize_batched_host_f64": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_ll, c_int, c_int, c_int, c_int, c_int]),
    "jpd_diagonalize_batched_host_f32": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_ll, c_int, c_int, c_int, c_int, c_int]),
    "jpd_diagonalize_batched_host_multi_f64": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_ll, c_int, c_int, c_int, c_int, c_int, c_vp]),
    "jpd_diagonalize_batched_host_multi_f32": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_ll, c_int, c_int, c_int, c_int, c_int, c_vp]),
    "jpd_diagonalize_batched_host_packed_f64": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_ll, c_int, c_int, c_int, c_int, c_int]),
    "jpd_diagonalize_batched_host_packed_f32": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_ll, c_int, c_int, c_int, c_int, c_int]),
    "jpd_host_register": (c_int, [c_vp, c_ull]),
    "jpd_host_unregister": (c_int, [c_vp]),
    "jpd_quaternion_from_correlation_host_f64": (c_int, [c_vp, c_vp, c_vp, c_vp, c_ll, c_int, c_int, c_int, c_int]),
    "jpd_quaternion_from_correlation_host_f32": (c_int, [c_vp, c_vp, c_vp, c_vp, c_ll, c_int, c_int, c_int, c_int]),
    "jpd_principal_axes_host_f64": (c_int, [c_vp, c_vp, c_vp, c_vp, c_ll, c_int, c_int, c_int]),
    "jpd_principal_axes_host_f32": (c_int, [c_vp, c_vp, c_vp, c_vp, c_ll, c_int, c_int, c_int]),
    "jpd_quaternion_from_correlation_f64": (c_int, [c_vp, c_vp, c_vp, c_vp, c_ll, c_int, c_int, c_int, c_vp]),
    "jpd_quaternion_from_correlation_f32": (c_int, [c_vp, c_vp, c_vp, c_vp, c_ll, c_int, c_int, c_int, c_vp]),
    "jpd_principal_axes_f64": (c_int, [c_vp, c_vp, c_vp, c_vp, c_ll, c_int, c_int, c_vp]),
    "jpd_principal_axes_f32": (c_int, [c_vp, c_vp, c_vp, c_vp, c_ll, c_int, c_int, c_vp]),
    "jpd_convert_layout_f64": (c_int, [c_vp, c_int, c_vp, c_int, c_int, c_ll, c_vp]),
    "jpd_convert_layout_f32": (c_int, [c_vp, c_int, c_vp, c_int, c_int, c_ll, c_vp]),
    "jpd_slice": (None, [c_ll, c_int, c_int, ctypes.POINTER(c_ll), ctypes.POINTER(c_ll)]),
    "jpd_fill_synthetic_f64": (c_int, [c_vp, c_int, c_ll, c_int, c_int, c_ull, c_ll, c_vp]),
    "jpd_fill_synthetic_f32": (c_int, [c_vp, c_int, c_ll, c_int, c_int, c_ull, c_ll, c_vp]),
    "jpd_fma_probe": (c_ll, [c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "jpd_algorithmic_bytes": (c_ll, [c_int, c_int, c_int]),
    "jpd_kernel_tier": (c_int, [c_int, c_int, c_int]),
    "jpd_launch_count": (c_ll, []),
    "jpd_release": (c_int, []),
    "jpd_error_string": (ctypes.c_char_p, [c_int]),
    "jpd_last_cuda_error": (ctypes.c_char_p, []),
    "jpd_version": (c_int, []),
}


class JpdError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile libjpd.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise JpdError("building libjpd.so failed (see output above)")
    return LIB_PATH


def lib() -> ctypes.CDLL:
    """The loaded library; raises (never falls back) when it is not there."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise JpdError(
                f"{LIB_PATH} not found: build it with `make -C jacobi_pd_b200/csrc` "
                "(there is no CPU or PyTorch fallback)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(code: int, what: str) -> None:
    if code != 0:
        l = lib()
        msg = l.jpd_error_string(code).decode()
        cuda = l.jpd_last_cuda_error().decode()
        raise JpdError(f"{what} failed: {msg} ({code})" + (f": {cuda}" if cuda else ""))
