"""-m gpu: the reference's OWN test program (tests/test_jacobi.cpp, compiled unchanged against
include/jacobi_pd.hpp of this repo by tests/dropin/Makefile) run on the GPU with the
reference CI's command lines (.circleci/config.yml:36-57) plus explicit seeds, in all four
container variants, and in coverage mode (n_tests = 0: sort-order asserts for three
criteria, DO_NOT_SORT, max_num_sweeps = 0; tests/test_jacobi.cpp:558-753)."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")

# n_size n_matr emin emax n_degeneracy n_tests seed
CI_MATRIX = [
    "2 1000 0.1 10 1 1 11", "2 1000 0.1 10 2 1 12",
    "5 1000 0.01 100.0 1 1 13", "5 1000 0.01 100.0 2 1 14",
    "20 100 1e-05 1e+05 1 1 15", "20 100 1e-05 1e+05 5 1 16",
    "50 10 1e-09 1e+09 1 1 17", "50 10 1e-09 1e+09 10 1 18",
]


def run(binary, args):
    path = os.path.join(REF_DIR, binary)
    # under -m gpu a missing binary is a FAILURE, not a skip: a box without the prebuilt drop-in
    # programs must not turn this row silently green
    assert os.path.exists(path), (f"{binary} not built: run __graft_entry__.build() (tests/dropin/Makefile) in the "
                                  "container that has /root/reference; the binary travels with gpurun")
    res = subprocess.run([path] + args.split(), capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "test passed" in res.stdout


@pytest.mark.parametrize("args", CI_MATRIX)
def test_reference_ci_matrix_on_gpu(args):
    run("test_jacobi_dropin", args)


@pytest.mark.parametrize("binary", ["test_jacobi_dropin", "test_jacobi_dropin_vec", "test_jacobi_dropin_arr",
                                    "test_jacobi_dropin_carr"])
def test_container_variants(binary):
    run(binary, "5 10 1e-04 1e+04 1 1 21")


@pytest.mark.parametrize("args", ["5 10 0.01 100 1 0 1", "3 20 0.1 10 2 0 2", "4 20 0 1 1 0 3", "12 5 0.01 100 3 0 4"])
def test_coverage_mode_sort_orders_and_failure_path(args):
    run("test_jacobi_dropin", args)


def test_cpp_batched_front_end():
    path = os.path.join(ROOT, "tests", "dropin", "test_batched_frontend")
    assert os.path.exists(path), "tests/dropin/test_batched_frontend not built (tests/dropin/Makefile)"
    res = subprocess.run([path], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "test passed" in res.stdout, res.stdout + res.stderr


def test_cpp_device_front_end_and_fused_producers():
    """CUDA C++ caller of BatchedJacobi::Diagonalize / QuaternionFromCorrelation / PrincipalAxes."""
    path = os.path.join(ROOT, "tests", "dropin", "test_device_frontend")
    assert os.path.exists(path), "tests/dropin/test_device_frontend not built (tests/dropin/Makefile)"
    res = subprocess.run([path], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "test passed" in res.stdout, res.stdout + res.stderr
