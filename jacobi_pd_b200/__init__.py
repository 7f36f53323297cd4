"""jacobi_pd_b200 -- B200-native batched Jacobi eigensolver behind jacobi_pd's Diagonalize API.

Product code: ``csrc/`` (CUDA kernels + C ABI -> ``libjpd.so``) and
``include/jacobi_pd_cuda.hpp`` (C++ front end).  This package is the Python binding of
the same C ABI.  The CPU oracle lives in ``oracle/`` and is never imported from here.
"""
from .api import (DO_NOT_SORT, SORT_DECREASING_EVALS, SORT_INCREASING_EVALS,
                  SORT_DECREASING_ABS_EVALS, SORT_INCREASING_ABS_EVALS, LAYOUT_AOS, LAYOUT_SOA,
                  Jacobi, diagonalize_batched, diagonalize_batched_host, fill_synthetic,
                  slice_of, algorithmic_bytes, quaternion_from_correlation, principal_axes, convert_layout, layout_shape,
                  quaternion_from_correlation_host, principal_axes_host, host_register, host_unregister,
                  LAYOUT_PACKED_AOS, LAYOUT_PACKED_SOA, LAYOUT_POINTERS)
from ._lib import JpdError, build, lib, LIB_PATH

__all__ = [
    "DO_NOT_SORT", "SORT_DECREASING_EVALS", "SORT_INCREASING_EVALS", "SORT_DECREASING_ABS_EVALS",
    "SORT_INCREASING_ABS_EVALS", "LAYOUT_AOS", "LAYOUT_SOA", "Jacobi", "diagonalize_batched",
    "diagonalize_batched_host", "fill_synthetic", "slice_of", "algorithmic_bytes",
    "quaternion_from_correlation", "principal_axes", "convert_layout", "layout_shape",
    "quaternion_from_correlation_host", "principal_axes_host", "host_register", "host_unregister",
    "LAYOUT_PACKED_AOS", "LAYOUT_PACKED_SOA", "LAYOUT_POINTERS", "JpdError",
    "build", "lib", "LIB_PATH",
]
