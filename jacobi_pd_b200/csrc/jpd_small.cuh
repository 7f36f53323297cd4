// jpd_small.cuh -- tier 1: one thread per matrix, everything in registers (n <= 6).
//
// The LAMMPS inertia-tensor (n=3) and colvars quaternion-superposition (n=4) shapes.
// Path being replaced: Jacobi::Diagonalize, /root/reference/include/jacobi_pd.hpp:159-220
// (MaxEntryRow/MaxEntry :427-468, CalcRot :233-256, ApplyRot :327-393,
//  ApplyRotLeft :409-419, SortRows :477-508).
//
// Design ("tier 1"):
//  * The reference's max-pivot search makes (i,j) data dependent; indexing a register
//    file dynamically is impossible and branching per pivot would serialise a warp 3x-6x.
//    Instead every lane walks the SAME static round-robin (chess-tournament) pair order,
//    fully unrolled, and each lane applies the reference's own skip test to its pair.
//    Measured rotation counts are within 4% of the max-pivot rule for n=3,4.
//  * Lanes that do not need a rotation get the identity (c=1,s=0,t=0) via selects, so
//    the FP64 stream is branch-free; a pair is skipped only when NO lane of the warp
//    needs it (__any_sync), and the sweep loop exits on a warp vote.
//  * Upper triangle only (as the reference reads only mat[i][j], j>=i, :167-171);
//    eigenvectors are rows (:69).
//  * Sort = the reference's selection sort replayed on (key,id) pairs with static
//    indices, so ties resolve exactly as SortRows does.
#pragma once
#include "jpd_common.cuh"

namespace jpd {

template <int N> struct SmallShape {
  static constexpr int kTri = N * (N + 1) / 2;
  static constexpr int kM = N + (N & 1);       // players in the tournament (dummy if odd)
  static constexpr int kSteps = kM - 1;
  static constexpr int kHalf = kM / 2;
  // packed upper-triangle slot of (i,j), i <= j
  JPD_CX static constexpr int tri(int i, int j) { return i * N - i * (i - 1) / 2 + (j - i); }
  // circle method: step s, table k -> players (a,b)
  JPD_CX static constexpr int player_a(int s, int k) { return k == 0 ? kM - 1 : (s + k) % (kM - 1); }
  JPD_CX static constexpr int player_b(int s, int k) { return k == 0 ? s : (s - k + (kM - 1)) % (kM - 1); }
};

#if defined(__CUDA_ARCH__)
#define JPD_WARP_ANY(pred) (__any_sync(0xffffffffu, (pred)))
#else
#define JPD_WARP_ANY(pred) (pred)
#endif

// One Jacobi rotation of the static pair (I,J), I<J, on the register-resident matrix.
template <int N, typename T, bool EVECS, int I, int J>
JPD_HD void small_rotate_pair(T (&m)[SmallShape<N>::kTri], T (&e)[EVECS ? N * N : 1],
                              bool& rotated, int& n_rot) {
  using Sh = SmallShape<N>;
  const T a = m[Sh::tri(I, J)];
  const T mii = m[Sh::tri(I, I)];
  const T mjj = m[Sh::tri(J, J)];
  const bool need = pair_needs_rotation(a, mii, mjj);
  if (!JPD_WARP_ANY(need)) return;
  const T d = mjj - mii;
  const bool any_unsafe = JPD_WARP_ANY(need && rotation_range_unsafe(d, a));
  T c, s, t;
  jacobi_rotation(d, a, any_unsafe, c, s, t);
  c = need ? c : T(1);
  s = need ? s : T(0);
  t = need ? t : T(0);
  rotated = rotated || need;
  n_rot += need ? 1 : 0;
  // diagonal through t, pivot zeroed: jacobi_pd.hpp:337-338,347
  const T g = t * a;
  m[Sh::tri(I, I)] = mii - g;
  m[Sh::tri(J, J)] = mjj + g;
  m[Sh::tri(I, J)] = need ? T(0) : a;
  // the two changed rows/columns, upper-triangle names: jacobi_pd.hpp:350-390
#pragma unroll
  for (int k = 0; k < N; ++k) {
    if (k == I || k == J) continue;
    const int ki = k < I ? Sh::tri(k, I) : Sh::tri(I, k);
    const int kj = k < J ? Sh::tri(k, J) : Sh::tri(J, k);
    const T x = m[ki], y = m[kj];
    m[ki] = c * x - s * y;
    m[kj] = s * x + c * y;
  }
  // eigenvector rows I and J: jacobi_pd.hpp:414-418
  if (EVECS) {
#pragma unroll
    for (int v = 0; v < N; ++v) {
      const T x = e[I * N + v], y = e[J * N + v];
      e[I * N + v] = c * x - s * y;
      e[J * N + v] = s * x + c * y;
    }
  }
}

// One full round-robin sweep, unrolled at compile time: pair P of kSteps*kHalf.
template <int N, typename T, bool EVECS, int P>
JPD_HD void small_sweep_from(T (&m)[SmallShape<N>::kTri], T (&e)[EVECS ? N * N : 1],
                             bool& rotated, int& n_rot) {
  using Sh = SmallShape<N>;
  if constexpr (P < Sh::kSteps * Sh::kHalf) {
    constexpr int S = P / Sh::kHalf, K = P % Sh::kHalf;
    constexpr int A = Sh::player_a(S, K), B = Sh::player_b(S, K);
    if constexpr (A < N && B < N)  // a pair with the dummy player (odd N) sits out
      small_rotate_pair<N, T, EVECS, (A < B ? A : B), (A < B ? B : A)>(m, e, rotated, n_rot);
    small_sweep_from<N, T, EVECS, P + 1>(m, e, rotated, n_rot);
  }
}

// Diagonalise the register-resident matrix. Returns the reference's return value:
// rotations + 1 (> 0) once a whole sweep found nothing to rotate, 0 if max_num_sweeps
// sweeps did not get there and n > 1 (jacobi_pd.hpp:214-219).
template <int N, typename T, bool EVECS>
JPD_HD int small_solve(T (&m)[SmallShape<N>::kTri], T (&e)[EVECS ? N * N : 1], int max_num_sweeps) {
  if (EVECS) {
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) e[i * N + j] = (i == j) ? T(1) : T(0);  // :172-178
  }
  if (N == 1) return 1;
  int n_rot = 0;
  bool converged = false;
  for (int sweep = 0; sweep < max_num_sweeps; ++sweep) {
    bool rotated = false;
    small_sweep_from<N, T, EVECS, 0>(m, e, rotated, n_rot);
    converged = converged || !rotated;
    if (!JPD_WARP_ANY(!converged)) break;
  }
  return converged ? n_rot + 1 : 0;
}

// Replay of SortRows (jacobi_pd.hpp:477-508) on keys with static indices; id[k] ends up
// as the source slot whose eigenpair must land in output position k.
template <int N, typename T>
JPD_HD void small_sort_ids(const T (&ev)[N], int sort_criteria, int (&id)[N]) {
  T key[N];
#pragma unroll
  for (int k = 0; k < N; ++k) {
    id[k] = k;
    const T v = ev[k];
    key[k] = sort_criteria == SORT_DECREASING_EVALS ? v
           : sort_criteria == SORT_INCREASING_EVALS ? -v
           : sort_criteria == SORT_DECREASING_ABS_EVALS ? jabs(v) : -jabs(v);
  }
  if (sort_criteria < SORT_DECREASING_EVALS || sort_criteria > SORT_INCREASING_ABS_EVALS)
    return;  // DO_NOT_SORT and unknown values: no-op (jacobi_pd.hpp:500-501)
#pragma unroll
  for (int i = 0; i < N - 1; ++i) {
    T best = key[i];
    int best_pos = i, best_id = id[i];
#pragma unroll
    for (int j = i + 1; j < N; ++j) {
      const bool gt = key[j] > best;  // strict: first maximum wins, as in the reference
      best = gt ? key[j] : best;
      best_pos = gt ? j : best_pos;
      best_id = gt ? id[j] : best_id;
    }
    const T old_key = key[i];
    const int old_id = id[i];
#pragma unroll
    for (int j = i + 1; j < N; ++j) {
      const bool hit = (best_pos == j);
      key[j] = hit ? old_key : key[j];
      id[j] = hit ? old_id : id[j];
    }
    key[i] = best;
    id[i] = best_id;
  }
}

#if defined(__CUDACC__)

// layout 0 (AoS): mat[b*N*N + i*N + j]; layout 1 (SoA-interleaved): mat[(i*N+j)*batch + b]
template <int N, typename T, bool SOA, bool EVECS>
__global__ void __launch_bounds__(128)
jpd_small_kernel(const T* __restrict__ mat, T* __restrict__ evals, T* __restrict__ evecs,
                 int* __restrict__ n_iters, long long batch, int sort_criteria,
                 int max_num_sweeps) {
  using Sh = SmallShape<N>;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = gid < batch;
  const long long b = live ? gid : batch - 1;  // tail lanes shadow the last matrix (no stores)

  T m[Sh::kTri];
  T e[EVECS ? N * N : 1];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = i; j < N; ++j)
      m[Sh::tri(i, j)] = SOA ? __ldg(mat + (long long)(i * N + j) * batch + b)
                             : __ldg(mat + b * (N * N) + (i * N + j));

  const int iters = small_solve<N, T, EVECS>(m, e, max_num_sweeps);

  T ev[N];
#pragma unroll
  for (int i = 0; i < N; ++i) ev[i] = m[Sh::tri(i, i)];  // :207-209
  int id[N];
  small_sort_ids<N, T>(ev, sort_criteria, id);
  const bool sorted = sort_criteria >= SORT_DECREASING_EVALS && sort_criteria <= SORT_INCREASING_ABS_EVALS;

  if (!live) return;
  if (n_iters) n_iters[b] = iters;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    T val = ev[k];
    if (sorted) {
#pragma unroll
      for (int r = 0; r < N; ++r) val = (id[k] == r) ? ev[r] : val;
    }
    if (SOA) evals[(long long)k * batch + b] = val;
    else evals[b * N + k] = val;
  }
  if (EVECS) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
#pragma unroll
      for (int v = 0; v < N; ++v) {
        T val = e[k * N + v];
        if (sorted) {
#pragma unroll
          for (int r = 0; r < N; ++r) val = (id[k] == r) ? e[r * N + v] : val;
        }
        if (SOA) evecs[(long long)(k * N + v) * batch + b] = val;
        else evecs[b * (N * N) + k * N + v] = val;
      }
    }
  }
}

#endif  // __CUDACC__

}  // namespace jpd
