// small_host_sim.cpp -- TEST INFRASTRUCTURE.  Compiles the tier-1 per-matrix routine
// (jacobi_pd_b200/csrc/jpd_small.cuh, the same template the CUDA kernel instantiates)
// for the HOST so that `pytest -m "not gpu"` can check its logic -- pair schedule, skip
// test, rotation algebra, scaling path, sort replay, return value -- without a GPU.
// It is never linked into libjpd.so and is not a CPU fallback of the product.
#include "jpd_small.cuh"

namespace {
template <int N, typename T>
void run(const T* mat, T* evals, T* evecs, int* iters, long long batch, int sort, int calc, int sweeps) {
  using Sh = jpd::SmallShape<N>;
  for (long long b = 0; b < batch; ++b) {
    auto load = [&](T (&mm)[Sh::kTri]) {
      for (int i = 0; i < N; ++i)
        for (int j = i; j < N; ++j) mm[Sh::tri(i, j)] = mat[b * N * N + i * N + j];
    };
    if (calc) {
      jpd::small_solve<N, T, true>(sweeps, load, [&](const T (&m)[Sh::kTri], const T (&e)[N * N], int it) {
        T ev[N]; int id[N];
        for (int i = 0; i < N; ++i) ev[i] = m[Sh::tri(i, i)];
        jpd::small_sort_ids<N, T>(ev, sort, id);
        for (int k = 0; k < N; ++k) {
          evals[b * N + k] = ev[id[k]];
          for (int v = 0; v < N; ++v) evecs[b * N * N + k * N + v] = e[id[k] * N + v];
        }
        if (iters) iters[b] = it;
      });
    } else {
      jpd::small_solve<N, T, false>(sweeps, load, [&](const T (&m)[Sh::kTri], const T (&)[1], int it) {
        T ev[N]; int id[N];
        for (int i = 0; i < N; ++i) ev[i] = m[Sh::tri(i, i)];
        jpd::small_sort_ids<N, T>(ev, sort, id);
        for (int k = 0; k < N; ++k) evals[b * N + k] = ev[id[k]];
        if (iters) iters[b] = it;
      });
    }
  }
}
template <typename T>
int dispatch(const T* mat, T* evals, T* evecs, int* iters, int n, long long batch, int sort, int calc, int sweeps) {
  switch (n) {
    case 1: run<1, T>(mat, evals, evecs, iters, batch, sort, calc, sweeps); return 0;
    case 2: run<2, T>(mat, evals, evecs, iters, batch, sort, calc, sweeps); return 0;
    case 3: run<3, T>(mat, evals, evecs, iters, batch, sort, calc, sweeps); return 0;
    case 4: run<4, T>(mat, evals, evecs, iters, batch, sort, calc, sweeps); return 0;
    case 5: run<5, T>(mat, evals, evecs, iters, batch, sort, calc, sweeps); return 0;
    case 6: run<6, T>(mat, evals, evecs, iters, batch, sort, calc, sweeps); return 0;
    default: return -1;
  }
}
}  // namespace

extern "C" {
int sim_small_f64(const double* mat, double* evals, double* evecs, int* iters, int n, long long batch,
                  int sort, int calc, int sweeps) {
  return dispatch<double>(mat, evals, evecs, iters, n, batch, sort, calc, sweeps);
}
int sim_small_f32(const float* mat, float* evals, float* evecs, int* iters, int n, long long batch,
                  int sort, int calc, int sweeps) {
  return dispatch<float>(mat, evals, evecs, iters, n, batch, sort, calc, sweeps);
}
}
