"""Comparator implementing the north-star tolerances (SURVEY 8(c) "comparator").

All inputs are AoS numpy arrays: mat (B,n,n), evals (B,n), evecs (B,n,n) with eigenvector
k in ROW k.  `ref_*` come from the CPU oracle / reference on the same inputs.

  eigenvalues   |l - l_ref| <= tol * max|l_ref|   at the SAME position after sorting (for the
                two ABS criteria: |l| by position and the signed values as a multiset)
  residual      max_k ||M v_k - l_k v_k||_inf <= tol * n * max|l|
  orthogonality max |V V^T - I| <= tol * n
  eigenvectors  rows equal up to sign where the reference spectrum is well separated
  n_iters       (gpu > 0) == (ref > 0)
with tol = 1e-12 (fp64) or 1e-5 (fp32).
"""
from __future__ import annotations

import numpy as np

TOL = {np.dtype("float64"): 1e-12, np.dtype("float32"): 1e-5}


def sym_from_upper(mat):
    up = np.triu(mat)
    return up + np.triu(mat, 1).swapaxes(-1, -2)


def scale_of(ref_evals):
    s = np.abs(ref_evals.astype(np.float64)).max(axis=1)
    return np.where(s > 0, s, 1.0)


def check_evals(evals, ref_evals, tol=None):
    tol = TOL[np.dtype(evals.dtype)] if tol is None else tol
    sc = scale_of(ref_evals)
    err = np.abs(evals.astype(np.float64) - ref_evals.astype(np.float64)).max(axis=1) / sc
    bad = np.flatnonzero(~(err <= tol))
    assert bad.size == 0, f"eigenvalue mismatch in {bad.size} matrices, worst rel err {np.nanmax(err):.3e} (tol {tol:g}), first bad {bad[:5]}"
    return float(err.max()) if err.size else 0.0


def check_residual_orth(mat, evals, evecs, tol=None):
    tol = TOL[np.dtype(evals.dtype)] if tol is None else tol
    n = mat.shape[-1]
    M = sym_from_upper(mat.astype(np.float64))
    V = evecs.astype(np.float64)
    L = evals.astype(np.float64)
    sc = np.maximum(np.abs(L).max(axis=1), np.abs(M).reshape(len(M), -1).max(axis=1))
    sc = np.where(sc > 0, sc, 1.0)
    R = np.einsum("bkv,bwv->bkw", V, M) - L[:, :, None] * V      # row k: (M v_k)^T - l_k v_k^T
    res = np.abs(R).reshape(len(M), -1).max(axis=1) / sc
    G = np.einsum("bkv,blv->bkl", V, V) - np.eye(n)
    orth = np.abs(G).reshape(len(M), -1).max(axis=1)
    bad = np.flatnonzero(~(res <= tol * n))
    assert bad.size == 0, f"residual too large in {bad.size} matrices, worst {np.nanmax(res):.3e} (tol {tol * n:g}), first bad {bad[:5]}"
    bad = np.flatnonzero(~(orth <= tol * n))
    assert bad.size == 0, f"eigenvectors not orthonormal in {bad.size} matrices, worst {np.nanmax(orth):.3e} (tol {tol * n:g})"
    return float(res.max()), float(orth.max())


def check_evecs_up_to_sign(evecs, ref_evecs, ref_evals, min_rel_gap=1e-6, factor=200.0):
    """Rows must agree up to sign wherever the reference eigenvalue is isolated: the
    allowed difference grows as eps*n/relative_gap (standard eigenvector conditioning)."""
    eps = np.finfo(evecs.dtype).eps
    n = evecs.shape[-1]
    L = ref_evals.astype(np.float64)
    sc = scale_of(ref_evals)[:, None]
    diff = np.abs(L[:, :, None] - L[:, None, :]) + np.eye(n)[None] * 1e300
    gap = diff.min(axis=2) / sc                                     # (B,n) relative gap of each eigenvalue
    V, W = evecs.astype(np.float64), ref_evecs.astype(np.float64)
    d = np.minimum(np.abs(V - W).max(axis=2), np.abs(V + W).max(axis=2))   # (B,n)
    allowed = factor * eps * n / np.maximum(gap, 1e-300)
    mask = gap >= min_rel_gap
    bad = mask & ~(d <= allowed)
    assert not bad.any(), f"eigenvectors differ beyond sign in {int(bad.sum())} rows, worst {d[mask].max():.3e}"
    return float(d[mask].max()) if mask.any() else 0.0, int(mask.sum())


def check_iters(iters, ref_iters):
    a, b = np.asarray(iters) > 0, np.asarray(ref_iters) > 0
    assert np.array_equal(a, b), f"convergence flags differ in {int((a != b).sum())} matrices"


def check_sorted(evals, sort_criteria):
    e = evals.astype(np.float64)
    if sort_criteria == 1:
        assert (np.diff(e, axis=1) <= 0).all()
    elif sort_criteria == 2:
        assert (np.diff(e, axis=1) >= 0).all()
    elif sort_criteria == 3:
        assert (np.diff(np.abs(e), axis=1) <= 0).all()
    elif sort_criteria == 4:
        assert (np.diff(np.abs(e), axis=1) >= 0).all()


def full_check(mat, out, ref, sort_criteria, calc_evecs=True, compare_vectors=True):
    evals, evecs, iters = out
    ref_evals, ref_evecs, ref_iters = ref[:3]
    stats = {}
    if sort_criteria in (3, 4):
        # ABS criteria: two eigenvalues of opposite sign whose magnitudes agree to within the
        # rounding of either implementation may legitimately come out in either order, so the
        # KEYS are compared by position and the signed values as a multiset
        stats["eval_err"] = check_evals(np.abs(evals), np.abs(ref_evals))
        check_evals(np.sort(evals, axis=1), np.sort(ref_evals, axis=1))
        check_sorted(evals, sort_criteria)
    elif sort_criteria != 0:
        stats["eval_err"] = check_evals(evals, ref_evals)
        check_sorted(evals, sort_criteria)
    else:  # unsorted order is implementation defined for n > 2: compare as multisets
        stats["eval_err"] = check_evals(np.sort(evals, axis=1), np.sort(ref_evals, axis=1))
    check_iters(iters, ref_iters)
    if calc_evecs:
        stats["res"], stats["orth"] = check_residual_orth(mat, evals, evecs)
        if compare_vectors and sort_criteria != 0:
            if sort_criteria in (3, 4):   # align by signed value first (see above)
                ia, ib = np.argsort(-evals, axis=1, kind="stable"), np.argsort(-ref_evals, axis=1, kind="stable")
                evecs = np.take_along_axis(evecs, ia[:, :, None], axis=1)
                ref_evecs = np.take_along_axis(ref_evecs, ib[:, :, None], axis=1)
                ref_evals = np.take_along_axis(ref_evals, ib, axis=1)
            stats["vec_diff"], stats["vec_rows"] = check_evecs_up_to_sign(evecs, ref_evecs, ref_evals)
    return stats
