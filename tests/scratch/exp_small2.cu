// dev experiment (not product): two matrices per thread, interleaved per pair, n=4/3 fp64 SoA
#include <cuda_runtime.h>
#include "jpd_small.cuh"
using namespace jpd;
#ifndef EXP_MINB
#define EXP_MINB 3
#endif
template <int N, typename T, bool SCALED, int P>
__device__ __forceinline__ void sweep2(T (&m0)[SmallShape<N>::kTri], T (&e0)[N*N], T (&m1)[SmallShape<N>::kTri], T (&e1)[N*N],
                                       int& r0, int& r1, bool& g0, bool& g1) {
  using Sh = SmallShape<N>;
  if constexpr (P < Sh::kSteps * Sh::kHalf) {
    constexpr int S = P / Sh::kHalf, K = P % Sh::kHalf;
    constexpr int A = Sh::player_a(S, K), B = Sh::player_b(S, K);
    if constexpr (A < N && B < N) {
      small_rotate_pair<N, T, true, SCALED, (A < B ? A : B), (A < B ? B : A)>(m0, e0, r0, g0);
      small_rotate_pair<N, T, true, SCALED, (A < B ? A : B), (A < B ? B : A)>(m1, e1, r1, g1);
    }
    sweep2<N, T, SCALED, P + 1>(m0, e0, m1, e1, r0, r1, g0, g1);
  }
}
template <int N>
__global__ void __launch_bounds__(128, EXP_MINB) k2(const double* __restrict__ mat, double* __restrict__ evals, double* __restrict__ evecs,
                                                  int* __restrict__ n_iters, long long batch, int sort, int max_sweeps) {
  using Sh = SmallShape<N>; using T = double;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long b = 2 * gid; if (b + 1 >= batch) b = batch - 2;
  T m0[Sh::kTri], m1[Sh::kTri], e0[N*N], e1[N*N];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = i; j < N; ++j) {
      const double2 v = __ldg(reinterpret_cast<const double2*>(mat + (long long)(i * N + j) * batch + b));
      m0[Sh::tri(i, j)] = v.x; m1[Sh::tri(i, j)] = v.y;
    }
#pragma unroll
  for (int i = 0; i < N * N; ++i) { e0[i] = (i % (N + 1) == 0) ? 1.0 : 0.0; e1[i] = e0[i]; }
  int r0 = 0, r1 = 0; bool c0 = false, c1 = false, rescale = false;
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    const bool a0 = small_status<N, T>(m0), a1 = small_status<N, T>(m1);
    c0 = c0 || !a0; c1 = c1 || !a1;
    if (!__any_sync(0xffffffffu, a0 | a1)) break;
    bool g0 = false, g1 = false;
    if (rescale) sweep2<N, T, true, 0>(m0, e0, m1, e1, r0, r1, g0, g1);
    else sweep2<N, T, false, 0>(m0, e0, m1, e1, r0, r1, g0, g1);
    rescale = rescale || __any_sync(0xffffffffu, g0 | g1);
  }
  if (2 * gid + 1 >= batch) return;
  // epilogue: two independent sorts + permuted stores (same as product kernel, vectorised)
  T ev0[N], ev1[N]; int id0[N], id1[N];
#pragma unroll
  for (int i = 0; i < N; ++i) { ev0[i] = m0[Sh::tri(i, i)]; ev1[i] = m1[Sh::tri(i, i)]; }
  small_sort_ids<N, T>(ev0, sort, id0); small_sort_ids<N, T>(ev1, sort, id1);
  reinterpret_cast<int2*>(n_iters + b)[0] = make_int2(c0 ? r0 + 1 : 0, c1 ? r1 + 1 : 0);
#pragma unroll
  for (int k = 0; k < N; ++k) {
    T x = ev0[k], y = ev1[k];
#pragma unroll
    for (int r = 0; r < N; ++r) { x = (id0[k] == r) ? ev0[r] : x; y = (id1[k] == r) ? ev1[r] : y; }
    reinterpret_cast<double2*>(evals + (long long)k * batch + b)[0] = make_double2(x, y);
  }
#pragma unroll
  for (int k = 0; k < N; ++k)
#pragma unroll
    for (int v = 0; v < N; ++v) {
      T x = e0[k * N + v], y = e1[k * N + v];
#pragma unroll
      for (int r = 0; r < N; ++r) { x = (id0[k] == r) ? e0[r * N + v] : x; y = (id1[k] == r) ? e1[r * N + v] : y; }
      reinterpret_cast<double2*>(evecs + (long long)(k * N + v) * batch + b)[0] = make_double2(x, y);
    }
}
extern "C" __attribute__((visibility("default"))) int exp_run(const double* mat, double* evals, double* evecs, int* iters, int n, long long batch, void* st) {
  const long long threads = (batch + 1) / 2; const unsigned blocks = (unsigned)((threads + 127) / 128);
  if (n == 4) k2<4><<<blocks, 128, 0, (cudaStream_t)st>>>(mat, evals, evecs, iters, batch, 1, 50);
  else if (n == 3) k2<3><<<blocks, 128, 0, (cudaStream_t)st>>>(mat, evals, evecs, iters, batch, 1, 50);
  else return -1;
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}
