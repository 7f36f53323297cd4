"""-m gpu: the fused producers against the CPU oracle.

colvars path: 3x3 correlation -> 4x4 overlap matrix S -> leading eigenpair (quaternion);
LAMMPS path: 6 inertia components -> principal moments + right-handed axes as columns.
The oracle diagonalises the explicitly built S / inertia tensor (Diagonalize,
/root/reference/include/jacobi_pd.hpp:159-220, SORT_DECREASING_EVALS); tolerances as in
tests/parity.py (1e-12 fp64, 1e-5 fp32, scaled by max|lambda|)."""
import numpy as np
import pytest

from oracle import binding as ob
import parity

pytestmark = pytest.mark.gpu


def overlap_from_correlation(R):
    xx, xy, xz, yx, yy, yz, zx, zy, zz = [R[:, i] for i in range(9)]
    S = np.zeros((len(R), 4, 4), R.dtype)
    S[:, 0, 0] = (xx + yy) + zz; S[:, 0, 1] = yz - zy; S[:, 0, 2] = zx - xz; S[:, 0, 3] = xy - yx
    S[:, 1, 1] = (xx - yy) - zz; S[:, 1, 2] = xy + yx; S[:, 1, 3] = zx + xz
    S[:, 2, 2] = (yy - xx) - zz; S[:, 2, 3] = yz + zy
    S[:, 3, 3] = (zz - xx) - yy
    return S + np.triu(S, 1).swapaxes(1, 2)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("layout", [0, 1])
def test_quaternion_from_correlation(dtype, layout):
    import torch
    import jacobi_pd_b200 as jpd
    rng = np.random.default_rng(11)
    B = 5000
    R = rng.uniform(-1, 1, (B, 9)).astype(dtype)
    R[0] = 0                                   # all-zero correlation: lambda 0, q = e0
    R[1] = np.eye(3).reshape(9)                # identity rotation: q = (1,0,0,0), lambda 3
    S = overlap_from_correlation(R)
    ref = ob.oracle_diagonalize(S, 1)
    tol = parity.TOL[np.dtype(dtype)]
    src = np.ascontiguousarray(R if layout == 0 else R.T)
    d = torch.from_numpy(src).cuda()
    lam, quat, it = jpd.quaternion_from_correlation(d, layout=layout)
    torch.cuda.synchronize()
    lam, quat, it = lam.cpu().numpy(), quat.cpu().numpy(), it.cpu().numpy()
    if layout == 1:
        quat = np.ascontiguousarray(quat.T)
    sc = parity.scale_of(ref[0])
    assert (np.abs(lam.astype(np.float64) - ref[0][:, 0]) <= tol * sc).all()
    parity.check_iters(it, ref[2])
    assert (quat[:, 0] >= 0).all()
    assert np.array_equal(quat[0], [1, 0, 0, 0]) and lam[0] == 0
    assert np.allclose(quat[1], [1, 0, 0, 0], atol=10 * tol) and abs(lam[1] - 3) <= 10 * tol
    # residual S q = lambda q and unit norm
    q = quat.astype(np.float64)
    res = np.abs(np.einsum("bij,bj->bi", S.astype(np.float64), q) - lam.astype(np.float64)[:, None] * q).max(axis=1)
    assert (res <= 4 * tol * np.maximum(sc, 1e-300)).all()
    assert (np.abs((q * q).sum(axis=1) - 1) <= 4 * tol).all()
    # equal to the oracle's leading row up to sign where the leading eigenvalue is isolated
    gap = (ref[0][:, 0] - ref[0][:, 1]) / sc
    ok = gap > 1e-3
    w = ref[1][:, 0, :].astype(np.float64)
    diff = np.minimum(np.abs(q - w).max(axis=1), np.abs(q + w).max(axis=1))
    assert (diff[ok] <= 200 * np.finfo(dtype).eps * 4 / gap[ok]).all()
    # all four eigenpairs: same contract as diagonalize_batched on S
    ev, vec, it2 = jpd.quaternion_from_correlation(d, layout=layout, all_eigenpairs=True)
    torch.cuda.synchronize()
    ev, vec = ev.cpu().numpy(), vec.cpu().numpy()
    if layout == 1:
        ev, vec = np.ascontiguousarray(ev.T), np.ascontiguousarray(vec.transpose(2, 0, 1))
    parity.full_check(S, (ev, vec, it2.cpu().numpy()), ref, 1)
    # and identical bits to the unfused path on the explicitly built S
    ev2, vec2, _ = jpd.diagonalize_batched(torch.from_numpy(S).cuda(), sort_criteria=1)
    assert np.array_equal(ev2.cpu().numpy(), ev) and np.array_equal(vec2.cpu().numpy(), vec)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("layout", [0, 1])
def test_principal_axes(dtype, layout):
    import torch
    import jacobi_pd_b200 as jpd
    rng = np.random.default_rng(12)
    B = 5000
    pts = rng.normal(size=(B, 12, 3)) * rng.uniform(0.2, 3.0, size=(B, 1, 3))
    r2 = (pts ** 2).sum(axis=2)
    I = np.einsum("bk,ij->bij", r2, np.eye(3)) - np.einsum("bki,bkj->bij", pts, pts)   # inertia tensors
    I = I.astype(dtype)
    I6 = np.stack([I[:, 0, 0], I[:, 1, 1], I[:, 2, 2], I[:, 1, 2], I[:, 0, 2], I[:, 0, 1]], axis=1)
    ref = ob.oracle_diagonalize(np.ascontiguousarray(I), 1)
    tol = parity.TOL[np.dtype(dtype)]
    src = np.ascontiguousarray(I6 if layout == 0 else I6.T)
    mom, axes, it = jpd.principal_axes(torch.from_numpy(src).cuda(), layout=layout)
    torch.cuda.synchronize()
    mom, axes, it = mom.cpu().numpy(), axes.cpu().numpy(), it.cpu().numpy()
    if layout == 1:
        mom, axes = np.ascontiguousarray(mom.T), np.ascontiguousarray(axes.transpose(2, 0, 1))
    parity.check_evals(mom, ref[0])
    parity.check_iters(it, ref[2])
    # columns are the eigenvectors: rows of axes^T satisfy the Diagonalize contract
    parity.check_residual_orth(I, mom, np.ascontiguousarray(axes.swapaxes(1, 2)))
    parity.check_evecs_up_to_sign(np.ascontiguousarray(axes.swapaxes(1, 2)), ref[1], ref[0])
    det = np.linalg.det(axes.astype(np.float64))
    assert (np.abs(det - 1) <= 100 * tol).all()          # right-handed frame
