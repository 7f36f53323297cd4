"""Generate the golden fixtures from the UNMODIFIED reference header (oracle/_ref).

Run in the build container (needs /root/reference to have been compiled into
oracle/_ref/libjpd_ref.so by `make -C oracle ref`):

    python tests/golden/make_fixtures.py

Writes
  tests/golden/known_answers.json    hand-picked inputs (README.md:67-69 example and the
                                     behaviours listed in SURVEY.md 8(c)) with the outputs the
                                     reference itself produced, as exact hex floats
  tests/golden/ref_fixtures_f64.npz  random + structured batches, all five SortCriteria,
  tests/golden/ref_fixtures_f32.npz  inputs and reference outputs (bit-exact arrays)
The fixtures travel with the repo; /root/reference does not exist on the GPU box.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import binding as ob  # noqa: E402


def hexlist(a):
    return [float(x).hex() for x in np.asarray(a, dtype=np.float64).ravel()]


def known_cases():
    c = []
    readme = [[2, 1, 1], [1, 2, -1], [1, -1, 2]]
    for sort in range(5):
        c.append(("readme_example_sort%d" % sort, readme, sort, True, 50))
    c.append(("two_by_two", [[1, 2], [2, 1]], 1, True, 50))
    c.append(("abs_tie_decreasing", [[0, 1], [1, 0]], 3, True, 50))
    c.append(("abs_tie_increasing", [[0, 1], [1, 0]], 4, True, 50))
    c.append(("already_diagonal", [[3, 0, 0], [0, 1, 0], [0, 0, 2]], 1, True, 50))
    c.append(("lower_triangle_ignored", [[2, 1, 1], [99, 2, -1], [99, 99, 2]], 1, True, 50))
    c.append(("all_zero", [[0, 0, 0], [0, 0, 0], [0, 0, 0]], 1, True, 50))
    c.append(("kappa_overflow", [[1e200, 1], [1, 0]], 1, True, 50))
    c.append(("tiny_and_denormal", [[1e-300, 1e-310], [1e-310, -1e-300]], 1, True, 50))
    wide = np.diag([1e20, 1e20, 1.0, 1.0]); wide[0, 1] = wide[1, 0] = 1e3; wide[2, 3] = wide[3, 2] = 0.5
    c.append(("wide_magnitude_early_exit", wide.tolist(), 1, True, 50))
    c.append(("n_equals_one", [[-7.5]], 1, True, 50))
    c.append(("zero_sweeps", [[2, 1, 1], [1, 2, -1], [1, -1, 2]], 1, True, 0))
    c.append(("zero_sweeps_diag", [[1, 0, 0], [0, 3, 0], [0, 0, 2]], 1, True, 0))
    c.append(("one_sweep_n2", [[1, 2], [2, 1]], 1, True, 1))
    c.append(("two_sweeps_n2", [[1, 2], [2, 1]], 1, True, 2))
    c.append(("repeated_diagonal", np.diag([2.0, -2.0, 2.0, -2.0, 0.0]).tolist(), 3, True, 50))
    c.append(("repeated_diagonal_inc_abs", np.diag([2.0, -2.0, 2.0, -2.0, 0.0]).tolist(), 4, True, 50))
    return c


def main():
    assert ob.ref_available(), "build oracle/_ref first (make -C oracle ref)"
    out = []
    for name, mat, sort, calc, sweeps in known_cases():
        m = np.array([mat], dtype=np.float64)
        ev, vec, it, _ = ob.ref_diagonalize(m, sort, calc, sweeps, 1)
        out.append({"name": name, "n": m.shape[1], "mat": hexlist(m), "sort_criteria": sort,
                    "calc_evecs": calc, "max_num_sweeps": sweeps, "ret": int(it[0]),
                    "evals": hexlist(ev), "evecs": hexlist(vec),
                    "evals_readable": [float(x) for x in ev.ravel()]})
    with open(os.path.join(HERE, "known_answers.json"), "w") as f:
        json.dump({"source": "outputs of /root/reference/include/jacobi_pd.hpp via oracle/_ref (g++ -O2, no FMA)",
                   "cases": out}, f, indent=1)

    rng = np.random.default_rng(20261019)
    for dt, tag in ((np.float64, "f64"), (np.float32, "f32")):
        arrays = {}
        k = 0
        for n in (2, 3, 4, 5, 8, 16, 20):
            for sort in range(5):
                B = 6 if n <= 8 else 3
                A = rng.uniform(-1, 1, (B, n, n))
                A = np.triu(A) + np.triu(A, 1).transpose(0, 2, 1)
                if sort in (3, 4):   # exact |.| ties and repeated values
                    A[0] = np.diag(np.resize([1.5, -1.5, 0.25], n))
                if sort == 1:        # log-uniform magnitudes 1e-9..1e9 (test_jacobi.cpp-style ranges)
                    A[1] *= 10.0 ** rng.uniform(-9, 9)
                A = np.ascontiguousarray(A.astype(dt))
                ev, vec, it, _ = ob.ref_diagonalize(A, sort, True, 50, 1)
                arrays[f"c{k}_mat"] = A; arrays[f"c{k}_sort"] = np.int32(sort)
                arrays[f"c{k}_evals"] = ev; arrays[f"c{k}_evecs"] = vec; arrays[f"c{k}_iters"] = it
                k += 1
        arrays["n_cases"] = np.int32(k)
        np.savez_compressed(os.path.join(HERE, f"ref_fixtures_{tag}.npz"), **arrays)
    print("fixtures written:", os.listdir(HERE))


if __name__ == "__main__":
    main()
