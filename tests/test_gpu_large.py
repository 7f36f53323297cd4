"""-m gpu: full oracle comparison at the BASELINE batch sizes the oracle finishes in seconds
("no large-batch parity"): configs[0] in full (1 Mi random 3x3, seed 1234, both
layouts) and a 1 Mi sample of configs[1] (quaternion 4x4, seed 4321); plus a bounded version
of the randomized soak (sizes x seeds x sort criteria x dtypes x layouts).  The rare paths of
tier 1 (rescaling pass, tie replay of the sort) fire at <~1e-6 per matrix: only batches of
this size exercise them on real data."""
import numpy as np
import pytest

from oracle import binding as ob
import parity
from gpu_util import run_gpu

pytestmark = pytest.mark.gpu

MI = 1 << 20


@pytest.mark.parametrize("layout", [0, 1])
def test_config0_one_mi_3x3_in_full(layout):
    A = ob.fill_synthetic(3, MI, 0, 0, 1234)
    ref = ob.oracle_diagonalize(A, 1)
    out = run_gpu(A, 1, layout=layout)
    st = parity.full_check(A, out, ref, 1)
    assert st["vec_rows"] > 3 * MI * 0.99          # nearly every eigenvector is compared up to sign


@pytest.mark.parametrize("layout", [0, 1])
def test_config1_one_mi_quaternion_sample(layout):
    A = ob.fill_synthetic(4, MI, 0, 1, 4321)
    ref = ob.oracle_diagonalize(A, 1)
    out = run_gpu(A, 1, layout=layout)
    st = parity.full_check(A, out, ref, 1)
    assert st["vec_rows"] > 4 * MI * 0.99
    assert (out[2] > 0).all()


def test_config1_host_entry_same_bits_as_device_entry():
    A = ob.fill_synthetic(4, MI // 4, 0, 1, 4321, b0=3 * MI)     # another slice of the same stream
    dev = run_gpu(A, 1)
    host = run_gpu(A, 1, host_api=True)
    for a, b in zip(dev, host):
        assert np.array_equal(a, b)


SOAK_SIZES = list(range(1, 41)) + [45, 48, 50, 63, 64, 65, 80, 96, 97, 100, 127, 128, 129, 150, 192, 200, 255, 256]


@pytest.mark.parametrize("rep", [0, 1])
def test_randomized_soak(rep):
    """tests/scratch/soak.py, bounded: 2 x 58 sizes x 2 dtypes = 232 cases per run, random input
    family / sort criterion / layout per case (seeded)."""
    rng = np.random.default_rng(2026 + rep)
    for n in SOAK_SIZES:
        B = 64 if n <= 8 else (24 if n <= 40 else (6 if n <= 128 else 3))
        for dtype in (np.float64, np.float32):
            kind = int(rng.integers(0, 3))
            if kind == 0:
                A = rng.uniform(-1, 1, (B, n, n))
            elif kind == 1:
                A = rng.normal(size=(B, n, n)) * 10.0 ** rng.uniform(-6, 6, (B, 1, 1))
            else:
                q, _ = np.linalg.qr(rng.normal(size=(B, n, n)))
                lam = np.round(rng.uniform(-3, 3, (B, n)))            # heavy exact degeneracy
                A = np.einsum("bki,bk,bkj->bij", q, lam, q)
            A = np.ascontiguousarray((np.triu(A) + np.triu(A, 1).swapaxes(1, 2)).astype(dtype))
            sort = int(rng.integers(0, 5)); layout = int(rng.integers(0, 2))
            ref = ob.oracle_diagonalize(A, sort)
            out = run_gpu(A, sort, layout=layout)
            where = dict(rep=rep, n=n, dtype=np.dtype(dtype).name, kind=kind, sort=sort, layout=layout)
            try:
                if kind == 2 and sort in (3, 4):     # |key| ties: order of +-l not defined
                    parity.check_evals(np.abs(out[0]), np.abs(ref[0]))
                    parity.check_residual_orth(A, out[0], out[1])
                    parity.check_iters(out[2], ref[2])
                else:
                    parity.full_check(A, out, ref, sort, compare_vectors=(kind != 2))
            except AssertionError as ex:
                raise AssertionError(f"{where}: {ex}") from ex
