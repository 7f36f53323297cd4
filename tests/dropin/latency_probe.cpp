// dev tool: per-call latency of the drop-in scalar class (batch = 1 through the host entry point)
#include <chrono>
#include <cstdio>
#include "jacobi_pd.hpp"
int main() {
  using namespace jacobi_pd;
  for (int n : {3, 4, 16, 32}) {
    double **M, **evecs;
    matrix_alloc_jpd::Alloc2D(n, n, &M);
    matrix_alloc_jpd::Alloc2D(n, n, &evecs);
    std::vector<double> evals(n);
    for (int i = 0; i < n; ++i) for (int j = i; j < n; ++j) M[i][j] = M[j][i] = 1.0 / (1 + i + j) + (i == j ? i : 0);
    Jacobi<double, double *, double **, double const *const *> solver(n);
    for (int k = 0; k < 50; ++k) solver.Diagonalize(M, evals.data(), evecs);   // warm-up (context, staging)
    const int calls = 2000;
    auto t0 = std::chrono::steady_clock::now();
    for (int k = 0; k < calls; ++k) solver.Diagonalize(M, evals.data(), evecs);
    auto t1 = std::chrono::steady_clock::now();
    std::printf("n=%d: %.1f us per Diagonalize call (evals[0]=%.6f)\n", n, std::chrono::duration<double, std::micro>(t1 - t0).count() / calls, evals[0]);
    matrix_alloc_jpd::Dealloc2D(&M);
    matrix_alloc_jpd::Dealloc2D(&evecs);
  }
  return 0;
}
