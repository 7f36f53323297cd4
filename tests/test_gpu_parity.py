"""-m gpu: CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Tolerances are the north star's: eigenvalues within 1e-12*max|l| (fp64) / 1e-5 (fp32) at
the same sorted position, eigenvectors equal up to sign where the spectrum is separated,
residual and orthogonality within tol*n, same convergence flag.  See tests/parity.py.
"""
import numpy as np
import pytest

from oracle import binding as ob
import parity
from gpu_util import run_gpu

pytestmark = pytest.mark.gpu

SIZES_SMALL = [1, 2, 3, 4, 5, 6]
SIZES_BLOCK = [7, 8, 9, 12, 16, 17, 20, 31, 32, 33, 50, 64]


def _batch_for(n):
    return 4096 if n <= 6 else (256 if n <= 20 else (64 if n <= 64 else 8))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", SIZES_SMALL + SIZES_BLOCK)
def test_parity_uniform(n, dtype):
    A = ob.fill_synthetic(n, _batch_for(n), seed=1234 + n, dtype=dtype)
    ref = ob.oracle_diagonalize(A, 1)
    for layout in (0, 1):
        out = run_gpu(A, 1, layout=layout)
        parity.full_check(A, out, ref, 1)


@pytest.mark.parametrize("n", [2, 3, 4, 5, 8, 20, 33, 130])
@pytest.mark.parametrize("sort", [0, 1, 2, 3, 4])
def test_all_sort_criteria(n, sort):
    A = ob.fill_synthetic(n, 512 if n <= 6 else 64, seed=77 + n)
    ref = ob.oracle_diagonalize(A, sort)
    out = run_gpu(A, sort)
    parity.full_check(A, out, ref, sort)


@pytest.mark.parametrize("n", [3, 4, 6, 16, 40, 160])
def test_no_evecs_and_sweep_limits(n):
    A = ob.fill_synthetic(n, 300, seed=5 + n)
    ref = ob.oracle_diagonalize(A, 1)
    ev, vec, it = run_gpu(A, 1, calc=False)
    assert vec is None
    parity.check_evals(ev, ref[0])
    parity.check_iters(it, ref[2])
    # max_num_sweeps = 0: reference returns 0, evals = sorted diagonal, evecs = permuted identity
    ref0 = ob.oracle_diagonalize(A, 1, True, 0)
    ev0, vec0, it0 = run_gpu(A, 1, sweeps=0)
    assert (it0 == 0).all() and (ref0[2] == 0).all()
    assert np.array_equal(ev0, ref0[0]) and np.array_equal(vec0, ref0[1])


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", [96, 128, 129, 200, 255, 256])
def test_parity_large(n, dtype):
    """tier 3 (CTA) up to 128, tier 4 (4-CTA cluster, M in distributed shared memory) up to 256;
    more matrices than resident clusters so that the persistent loop is exercised"""
    A = ob.fill_synthetic(n, 5 if n <= 128 else 36, seed=9000 + n, dtype=dtype)
    ref = ob.oracle_diagonalize(A, 1)
    for layout in (0, 1):
        out = run_gpu(A, 1, layout=layout)
        parity.full_check(A, out, ref, 1)
    out2 = run_gpu(A, 1)                       # repeatable bit for bit (no cross-CTA race)
    for x, y in zip(out, out2):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", [33, 41, 46, 47, 50, 57, 63, 65, 81, 97, 98, 99, 111, 127])
def test_parity_tier3_both_kernels(n, dtype):
    """tier 3 dispatches between the first-generation kernel (2x2 blocks) and the block-ordering kernel
    (4x4 register tiles, scaled eigenvector rotations) by size and precision: sizes on both sides of
    every switch point, every padding (n mod 4), all sort criteria on one of them"""
    A = ob.fill_synthetic(n, 12, seed=4000 + n, dtype=dtype)
    ref = ob.oracle_diagonalize(A, 1)
    for layout in (0, 1):
        out = run_gpu(A, 1, layout=layout)
        parity.full_check(A, out, ref, 1)
    if n in (57, 99):
        for sort in (0, 2, 3, 4):
            parity.full_check(A, run_gpu(A, sort), ob.oracle_diagonalize(A, sort), sort)
        ev, vec, it = run_gpu(A, 1, calc=False)
        assert vec is None
        parity.check_evals(ev, ref[0])
        ref0 = ob.oracle_diagonalize(A, 1, True, 0)           # max_num_sweeps = 0
        ev0, vec0, it0 = run_gpu(A, 1, sweeps=0)
        assert (it0 == 0).all() and np.array_equal(ev0, ref0[0]) and np.array_equal(vec0, ref0[1])
        # repeatable bit for bit (shared-memory hazards between the pivot, tile and eigenvector phases
        # would show up as run-to-run differences); a larger batch so that the persistent loop reuses its buffers
        big = ob.fill_synthetic(n, 600, seed=4100 + n, dtype=dtype)
        first = run_gpu(big, 1)
        for _ in range(4):
            again = run_gpu(big, 1)
            assert all(np.array_equal(x, y) for x, y in zip(first, again))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", [131, 145, 175, 176, 177, 191, 193, 230, 253])
def test_parity_tier4_both_kernels(n, dtype):
    """tier 4 runs the block-ordering cluster kernel (jpd_cluster4.cuh: 4x4 register tiles, tile rows dealt to the
    four CTAs in groups of eight, pattern-B tiles that reach into a neighbour CTA, pivots exchanged as bulk copies
    on mbarriers): every padding, tile rows on either side of a group boundary, all sort criteria and the
    no-eigenvector path on two of them, bit-repeatability (the name dates from the size-dependent dispatch of the
    two tier-4 kernels; the first-generation one is covered by test_tier4_first_generation_kernel)"""
    A = ob.fill_synthetic(n, 5, seed=5000 + n, dtype=dtype)
    ref = ob.oracle_diagonalize(A, 1)
    for layout in (0, 1):
        out = run_gpu(A, 1, layout=layout)
        parity.full_check(A, out, ref, 1)
    if n in (177, 253):
        for sort in (0, 2, 3, 4):
            parity.full_check(A, run_gpu(A, sort), ob.oracle_diagonalize(A, sort), sort)
        ev, vec, it = run_gpu(A, 1, calc=False)
        assert vec is None
        parity.check_evals(ev, ref[0])
        ref0 = ob.oracle_diagonalize(A, 1, True, 0)           # max_num_sweeps = 0
        ev0, vec0, it0 = run_gpu(A, 1, sweeps=0)
        assert (it0 == 0).all() and np.array_equal(ev0, ref0[0]) and np.array_equal(vec0, ref0[1])
        big = ob.fill_synthetic(n, 70, seed=5100 + n, dtype=dtype)   # more matrices than resident clusters
        first = run_gpu(big, 1)
        for _ in range(2):
            again = run_gpu(big, 1)
            assert all(np.array_equal(x, y) for x, y in zip(first, again))
        parity.full_check(big, first, ob.oracle_diagonalize(big, 1), 1)


def test_tier4_first_generation_kernel():
    """JPD_TIER4_BLOCK=0 selects the first-generation cluster kernel (jpd_cluster.cuh); the switch is read once per
    process, hence the child process"""
    import os, subprocess, sys
    code = (
        "import sys; sys.path.insert(0, 'tests'); sys.path.insert(0, '.')\n"
        "import numpy as np, parity\n"
        "from oracle import binding as ob\n"
        "from gpu_util import run_gpu\n"
        "for n, dt in ((200, np.float64), (131, np.float32)):\n"
        "    A = ob.fill_synthetic(n, 5, seed=5200 + n, dtype=dt)\n"
        "    parity.full_check(A, run_gpu(A, 1), ob.oracle_diagonalize(A, 1), 1)\n"
        "print('ok')\n")
    env = dict(os.environ, JPD_TIER4_BLOCK="0")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "ok" in res.stdout, res.stdout + res.stderr


def _random_symmetric(n, batch, dtype, seed):
    rng = np.random.default_rng(seed)            # (the counter generator stops at n = 256)
    A = rng.uniform(-1, 1, (batch, n, n))
    return np.ascontiguousarray((np.triu(A) + np.triu(A, 1).swapaxes(1, 2)).astype(dtype))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", [257, 300])
def test_parity_block_jacobi_tier_vs_oracle(n, dtype):
    """tier 5 (block Jacobi over the whole device, jpd_blockjacobi.cuh) at sizes the oracle still
    finishes in seconds: all sort criteria, both layouts, calc_evecs off, max_num_sweeps = 0"""
    A = _random_symmetric(n, 3, dtype, 9000 + n)
    for sort in (1, 2, 3, 4, 0):
        ref = ob.oracle_diagonalize(A, sort)
        out = run_gpu(A, sort, layout=sort & 1)
        parity.full_check(A, out, ref, sort)
    ev, vec, it = run_gpu(A, 1, calc=False)
    assert vec is None
    parity.check_evals(ev, ob.oracle_diagonalize(A, 1)[0])
    ref0 = ob.oracle_diagonalize(A, 1, True, 0)
    ev0, vec0, it0 = run_gpu(A, 1, sweeps=0)
    assert (it0 == 0).all() and np.array_equal(ev0, ref0[0]) and np.array_equal(vec0, ref0[1])
    a, b = run_gpu(A, 1), run_gpu(A[:1], 1)      # a matrix's result does not depend on its batch
    assert np.array_equal(a[0][:1], b[0]) and np.array_equal(a[1][:1], b[1])


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("n", [384, 512, 1000])
def test_block_jacobi_tier_large(n, layout, dtype):
    """The sizes the reference benchmarks beyond the configs (benchmarks/README.md:19-20).  The
    oracle needs minutes per matrix here, so eigenvalues are checked against LAPACK
    (numpy.linalg.eigvalsh, fp64) with the north-star tolerance, eigenvectors through the
    size-independent properties: residual, orthonormality, order."""
    A = _random_symmetric(n, 2, dtype, 7000 + n)
    ev, vec, it = run_gpu(A, 1, layout=layout)
    assert (it > 0).all()
    lap = np.linalg.eigvalsh(A.astype(np.float64))[:, ::-1]
    # north-star tolerance (1e-12 fp64 / 1e-5 fp32, stated for the configs' n <= 256), but never
    # below what rounding alone allows at this size: n * eps / 2 (fp32, n = 1000: 3e-5)
    tol = max(parity.TOL[np.dtype(dtype)], 0.5 * n * np.finfo(dtype).eps)
    parity.check_evals(ev, lap.astype(dtype), tol=tol)
    parity.check_sorted(ev, 1)
    parity.check_residual_orth(A, ev, vec)


def test_host_api_matches_device_api():
    for n, layout in ((4, 0), (4, 1), (3, 0), (12, 0), (12, 1)):
        A = ob.fill_synthetic(n, 1000, seed=31 + n)
        a = run_gpu(A, 1, layout=layout)
        b = run_gpu(A, 1, layout=layout, host_api=True)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)


def test_synthetic_generator_bits_match_host_twin():
    import torch
    import jacobi_pd_b200 as jpd
    for n, kind, layout in ((3, 0, 0), (3, 0, 1), (4, 1, 0), (4, 1, 1), (32, 0, 0)):
        B = 777
        shape = (B, n, n) if layout == 0 else (n, n, B)
        for dt, tdt in ((np.float64, torch.float64), (np.float32, torch.float32)):
            d = torch.empty(shape, dtype=tdt, device="cuda")
            jpd.fill_synthetic(d, n, B, layout, kind, 4321, b0=5)
            h = ob.fill_synthetic(n, B, layout, kind, 4321, b0=5, dtype=dt)
            assert np.array_equal(d.cpu().numpy(), h)


@pytest.mark.parametrize("layout", [0, 1])
def test_multi_gpu_host_entry_slices_without_changing_results(layout):
    """jpd_diagonalize_batched_host_multi_*: the batch is cut into contiguous slices, one per
    device (all visible devices, at most 4); results must be bit-identical to the
    single-device call (independent units, no collective)."""
    import torch
    import jacobi_pd_b200 as jpd
    n_dev = max(1, min(4, torch.cuda.device_count()))
    for n, B in ((4, 10007), (12, 1001), (40, 67)):
        A = ob.fill_synthetic(n, B, seed=500 + n)
        src = np.ascontiguousarray(A if layout == 0 else A.transpose(1, 2, 0))
        es, vs = jpd.api.out_shapes(n, B, layout)
        outs = []
        for devices in (1, n_dev):
            ev, vec, it = np.empty(es), np.empty(vs), np.empty(B, np.int32)
            jpd.diagonalize_batched_host(src, ev, vec, it, n, B, layout, 1, True, 50, n_devices=devices) if devices > 1 else \
                jpd.diagonalize_batched_host(src, ev, vec, it, n, B, layout, 1, True, 50)
            outs.append((ev, vec, it))
        # also the multi entry with a single device
        ev, vec, it = np.empty(es), np.empty(vs), np.empty(B, np.int32)
        l = jpd.lib()
        rc = l.jpd_diagonalize_batched_host_multi_f64(src.ctypes.data, ev.ctypes.data, vec.ctypes.data, it.ctypes.data,
                                                      n, B, layout, 1, 1, 50, 1, None)
        assert rc == 0
        outs.append((ev, vec, it))
        for o in outs[1:]:
            for x, y in zip(outs[0], o):
                assert np.array_equal(x, y)
