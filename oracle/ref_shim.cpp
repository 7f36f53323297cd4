// ref_shim.cpp -- TEST INFRASTRUCTURE.  Thin C-ABI wrapper around the UNMODIFIED
// reference header.  It contains no algorithm of its own: it #includes
//   <jacobi_pd.hpp>   (found with -I/root/reference/include by oracle/Makefile)
// and loops jacobi_pd::Jacobi<S,S*,S**,S const*const*>::Diagonalize over an AoS
// batch, one Jacobi instance per OpenMP thread (the pattern SURVEY 8(b)/8(d)
// prescribes for the CPU baseline).  Output: oracle/_ref/libjpd_ref.so (git-ignored,
// travels to the GPU box with the snapshot; /root/reference itself does not).
//
// Used (a) to pin oracle/jpd_oracle.c bit-for-bit, (b) to generate tests/golden
// fixtures, (c) as bench.py's cpu_baseline / --impl reference arm (kind
// "reference").  Never loaded by the product path.
#include <cstddef>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "jacobi_pd.hpp"

namespace {

template <typename S>
int run_batch(const S *mat, S *evals, S *evecs, int *n_iters, int n, long long batch,
              int sort_criteria, int calc_evecs, int max_num_sweeps, int n_threads) {
  using Solver = jacobi_pd::Jacobi<S, S *, S **, S const *const *>;
  if (n <= 0 || batch <= 0) return 0;
  int used = 1;
#ifdef _OPENMP
  if (n_threads <= 0) n_threads = omp_get_max_threads();
  used = n_threads;
#pragma omp parallel num_threads(n_threads)
#endif
  {
    Solver solver(n);
    std::vector<const S *> in_rows(n);
    std::vector<S *> out_rows(n);
    std::vector<S> scratch((std::size_t)n * n);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (long long b = 0; b < batch; ++b) {
      const S *m = mat + (std::size_t)b * n * n;
      S *e = (calc_evecs && evecs) ? evecs + (std::size_t)b * n * n : scratch.data();
      for (int i = 0; i < n; ++i) {
        in_rows[i] = m + (std::size_t)i * n;
        out_rows[i] = e + (std::size_t)i * n;
      }
      int r = solver.Diagonalize(in_rows.data(), evals + (std::size_t)b * n, out_rows.data(),
                                 static_cast<typename Solver::SortCriteria>(sort_criteria),
                                 calc_evecs != 0, max_num_sweeps);
      if (n_iters) n_iters[b] = r;
    }
  }
  return used;
}

}  // namespace

extern "C" {

int jpdref_diagonalize_batched_f64(const double *mat, double *evals, double *evecs, int *n_iters,
                                   int n, long long batch, int sort_criteria, int calc_evecs,
                                   int max_num_sweeps, int n_threads) {
  return run_batch<double>(mat, evals, evecs, n_iters, n, batch, sort_criteria, calc_evecs,
                           max_num_sweeps, n_threads);
}

int jpdref_diagonalize_batched_f32(const float *mat, float *evals, float *evecs, int *n_iters,
                                   int n, long long batch, int sort_criteria, int calc_evecs,
                                   int max_num_sweeps, int n_threads) {
  return run_batch<float>(mat, evals, evecs, n_iters, n, batch, sort_criteria, calc_evecs,
                          max_num_sweeps, n_threads);
}

int jpdref_host_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
