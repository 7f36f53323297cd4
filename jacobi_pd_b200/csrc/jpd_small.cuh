// jpd_small.cuh -- tier 1: one thread per matrix, everything in registers (n <= 6).
//
// The LAMMPS inertia-tensor (n=3) and colvars quaternion-superposition (n=4) shapes.
// Path being replaced: Jacobi::Diagonalize, /root/reference/include/jacobi_pd.hpp:159-220
// (MaxEntryRow/MaxEntry :427-468, CalcRot :233-256, ApplyRot :327-393,
//  ApplyRotLeft :409-419, SortRows :477-508).
//
// Design ("tier 1"):
//  * The reference's max-pivot search makes (i,j) data dependent; a register file cannot
//    be indexed dynamically and branching per pivot would serialise a warp 3x-6x.
//    Instead every lane walks the SAME static round-robin (chess-tournament) pair order,
//    fully unrolled, and each lane applies the reference's own skip test to its pair.
//    Measured rotation counts are within 4% of the max-pivot rule for n=3,4.
//  * A sweep is ONE branch-free basic block: lanes whose pair needs no rotation get the
//    exact identity (c=1,s=0,t=0) through selects.  That lets ptxas interleave the
//    independent rotations of a round-robin step (FP64 latency hiding) and removes the
//    per-pair votes / register shuffles that branchy code needs (r1a profile: 64% of the
//    issued instructions were not FP64).
//  * Control flow is per sweep and warp-uniform: before a sweep every lane evaluates the
//    skip test on all its pairs; the warp leaves when no lane has work (__any_sync).
//  * Range safety without per-pair cost: the common sweep computes x = d^2+(2a)^2
//    directly; a warp votes once per sweep whether any lane holds entries so large/small
//    that x could overflow/underflow and only then runs the sweep variant that rescales
//    (d,a) by a power of two per pair.  A last-resort guard skips (for this sweep only) a
//    pair whose x still underflowed.
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

// exponent window in which the un-rescaled rotation is safe for every pair of a sweep
template <typename T> struct SmallRange {
  static constexpr int kLo = Fp<T>::kBias - (Fp<T>::kSafe * 5) / 6;   // entries >= 2^-400 (fp64)
  static constexpr int kHi = Fp<T>::kBias + Fp<T>::kSafe;             // entries <  2^481
  static constexpr int kXMin = Fp<T>::kBias - 2 * Fp<T>::kSafe + Fp<T>::kSafe / 8;  // x >= 2^-900
  static constexpr int kXMax = Fp<T>::kBias + 2 * Fp<T>::kSafe + Fp<T>::kSafe / 8;  // x <  2^1020
};

// |x| as an ordered integer: the high word of a double (the whole word of a float) with
// the sign bit cleared grows monotonically with |x|, so magnitudes can be compared and
// scaled by powers of two on the integer pipe.
JPD_HD int abs_key(double x) {
#if defined(__CUDA_ARCH__)
  return __double2hiint(x) & 0x7fffffff;
#else
  std::uint64_t u; std::memcpy(&u, &x, 8);
  return (int)((u >> 32) & 0x7fffffffu);
#endif
}
JPD_HD int abs_key(float x) {
#if defined(__CUDA_ARCH__)
  return __float_as_int(x) & 0x7fffffff;
#else
  std::uint32_t u; std::memcpy(&u, &x, 4);
  return (int)(u & 0x7fffffffu);
#endif
}
template <typename T> struct KeyShift;   // bit position of the exponent field inside abs_key
template <> struct KeyShift<double> { static constexpr int value = 20; };
template <> struct KeyShift<float>  { static constexpr int value = 23; };

// Skip test of the reference (jacobi_pd.hpp:191: a is below half an ulp of BOTH diagonal
// entries) evaluated as  |a| * 2^(kMant+2) <= min(|mii|,|mjj|)  on abs_key()s.  The bound is
// inside the reference's skip region, so a pair the reference would rotate is never skipped.
template <typename T>
JPD_HD bool key_needs_rotation(int ka, bool a_is_zero, int kii, int kjj) {
  const int kmin = kii < kjj ? kii : kjj;
  // (kmin - shift rather than ka + shift: the latter overflows int for |a| > 2^970)
  const int lowered = kmin - ((Fp<T>::kMantBits + 2) << KeyShift<T>::value);
  return !a_is_zero & (ka > lowered);
}

// Rows I and J of E, column W onwards, using what is known about the entries at compile time.
template <int N, typename T, int I, int J, unsigned long long NZ, unsigned long long ONE, int W>
JPD_HD void small_rotate_rows(T (&e)[N * N], T c, T s) {
  if constexpr (W < N) {
    constexpr bool xi = (NZ >> (I * N + W)) & 1ull, yj = (NZ >> (J * N + W)) & 1ull;
    constexpr bool x1 = (ONE >> (I * N + W)) & 1ull, y1 = (ONE >> (J * N + W)) & 1ull;
    if constexpr (xi && yj) {
      const T xx = e[I * N + W], yy = e[J * N + W];
      e[I * N + W] = jfma(c, xx, -(s * yy));
      e[J * N + W] = jfma(s, xx, c * yy);
    } else if constexpr (xi) {          // E_J entry is 0
      if constexpr (x1) { e[I * N + W] = c; e[J * N + W] = s; }
      else { const T xx = e[I * N + W]; e[I * N + W] = c * xx; e[J * N + W] = s * xx; }
    } else if constexpr (yj) {          // E_I entry is 0
      if constexpr (y1) { e[I * N + W] = -s; e[J * N + W] = c; }
      else { const T yy = e[J * N + W]; e[I * N + W] = -(s * yy); e[J * N + W] = c * yy; }
    }
    small_rotate_rows<N, T, I, J, NZ, ONE, W + 1>(e, c, s);
  }
}

// One Jacobi rotation of the static pair (I,J), I<J, branch-free.  The skip test is taken
// on the CURRENT entries (re-using the sweep-start flags instead was measured: stale flags
// cost ~4% more rotations and a longer warp-max sweep count).
//
// NZ / ONE: compile-time knowledge about E during the FIRST sweep, which starts from the
// identity (bit i*N+w set = entry (i,w) may be nonzero / is exactly 1).  Rows that are still
// unit vectors or sparse get their rotation with 0 or 2 multiplications per column instead
// of 4 FP64 instructions (n = 4: 48 instead of 96 for the sweep).  Later sweeps pass
// kDenseMask and compile to the plain update.
constexpr unsigned long long kDenseMask = ~0ull;
template <int N> JPD_CX constexpr unsigned long long identity_mask() {
  unsigned long long mk = 0;
  for (int i = 0; i < N; ++i) mk |= 1ull << (i * N + i);
  return mk;
}
// masks after rows I and J have been combined
template <int N> JPD_CX constexpr unsigned long long mask_after_nz(unsigned long long nz, int I, int J) {
  for (int w = 0; w < N; ++w) {
    const unsigned long long any = ((nz >> (I * N + w)) | (nz >> (J * N + w))) & 1ull;
    nz = (nz & ~((1ull << (I * N + w)) | (1ull << (J * N + w)))) | (any << (I * N + w)) | (any << (J * N + w));
  }
  return nz;
}
template <int N> JPD_CX constexpr unsigned long long mask_after_one(unsigned long long one, int I, int J) {
  for (int w = 0; w < N; ++w) one &= ~((1ull << (I * N + w)) | (1ull << (J * N + w)));
  return one;
}

template <int N, typename T, bool EVECS, bool SCALED, int I, int J, unsigned long long NZ, unsigned long long ONE>
JPD_HD void small_rotate_pair(T (&m)[SmallShape<N>::kTri], T (&e)[EVECS ? N * N : 1],
                              int& n_rot, bool& guard_hit) {
  using Sh = SmallShape<N>;
  const T a = m[Sh::tri(I, J)];
  const T mii = m[Sh::tri(I, I)];
  const T mjj = m[Sh::tri(J, J)];
  const bool flagged = key_needs_rotation<T>(abs_key(a), is_zero(a), abs_key(mii), abs_key(mjj));
  T d = mjj - mii;
  T as = a;
  // sign(t) = sign(a) sign(d); the reference takes t = +1 when d == 0 whatever the sign of
  // a (jacobi_pd.hpp:237-239) -- kept, so that n == 2 agrees with it even unsorted.
  const T sign_src = is_zero(d) ? a : d;
  if (SCALED) rescale_for_rotation(d, as);
  // ---- rotation_from_scaled() of jpd_common.cuh, inlined so the guard can look at x
  const T a2 = as + as;
  const T x = jfma(a2, a2, d * d);
  // x == 0 (pair already exactly diagonal), or x under/overflowed in the un-rescaled
  // variant: leave the pair alone in this sweep; small_solve() then switches the warp to
  // the rescaling variant, which has no such limit.
  const int kx = abs_key(x);
  const bool x_ok = SCALED ? (kx != 0)
                           : ((unsigned)(kx - (SmallRange<T>::kXMin << KeyShift<T>::value)) <
                              (unsigned)((SmallRange<T>::kXMax - SmallRange<T>::kXMin) << KeyShift<T>::value));
  const bool need = flagged & x_ok;
  guard_hit = guard_hit | (flagged & !x_ok);
#if defined(JPD_SELECT_CST)
  const T p = rsqrt_refined(x);
  const T rho = jabs(d) * p;
  const T ap = xor_sign(as * p, sign_src);         // sin(2 theta) / 2 with the sign of t
  const T v = jfma(half_const(T(0)), rho, T(0.5));  // cos^2(theta)
  const T u = rsqrt_refined(v);
  T c = v * u;
  T s = ap * u;
  T t = s * u;
  c = need ? c : T(1);
  s = need ? s : T(0);
  t = need ? t : T(0);
#else
  // A pair that is left alone gets the exact identity through TWO selects: p = 0 makes
  // rho = 0, v = 1/2 (finite whatever x was), s = t = (a * 0) * u = 0, and c is replaced by 1.
  T p = rsqrt_refined(x);
  p = need ? p : T(0);
  const T rho = jabs(d) * p;
  const T ap = xor_sign(as * p, sign_src);         // sin(2 theta) / 2 with the sign of t
  const T v = jfma(half_const(T(0)), rho, T(0.5));  // cos^2(theta)
  const T u = rsqrt_refined(v);
  T c = v * u;
  const T s = ap * u;
  const T t = s * u;
  c = need ? c : T(1);
#endif
  n_rot += need ? 1 : 0;
  // diagonal through t, pivot zeroed: jacobi_pd.hpp:337-338,347
  m[Sh::tri(I, I)] = jfma(-t, a, mii);
  m[Sh::tri(J, J)] = jfma(t, a, mjj);
  m[Sh::tri(I, J)] = need ? T(0) : a;
  // the two changed rows/columns, upper-triangle names: jacobi_pd.hpp:350-390
#pragma unroll
  for (int k = 0; k < N; ++k) {
    if (k == I || k == J) continue;
    const int ki = k < I ? Sh::tri(k, I) : Sh::tri(I, k);
    const int kj = k < J ? Sh::tri(k, J) : Sh::tri(J, k);
    const T xx = m[ki], yy = m[kj];
    m[ki] = jfma(c, xx, -(s * yy));
    m[kj] = jfma(s, xx, c * yy);
  }
  // eigenvector rows I and J: jacobi_pd.hpp:414-418
  if constexpr (EVECS) small_rotate_rows<N, T, I, J, NZ, ONE, 0>(e, c, s);
}

// One full round-robin sweep, unrolled at compile time: pair P of kSteps*kHalf.
template <int N, typename T, bool EVECS, bool SCALED, int P, unsigned long long NZ = kDenseMask,
          unsigned long long ONE = 0ull>
JPD_HD void small_sweep_from(T (&m)[SmallShape<N>::kTri], T (&e)[EVECS ? N * N : 1],
                             int& n_rot, bool& guard_hit) {
  using Sh = SmallShape<N>;
  if constexpr (P < Sh::kSteps * Sh::kHalf) {
    constexpr int S = P / Sh::kHalf, K = P % Sh::kHalf;
    constexpr int A = Sh::player_a(S, K), B = Sh::player_b(S, K);
    if constexpr (A < N && B < N) {  // a pair with the dummy player (odd N) sits out
      constexpr int I = A < B ? A : B, J = A < B ? B : A;
      small_rotate_pair<N, T, EVECS, SCALED, I, J, NZ, ONE>(m, e, n_rot, guard_hit);
      small_sweep_from<N, T, EVECS, SCALED, P + 1, mask_after_nz<N>(NZ, I, J), mask_after_one<N>(ONE, I, J)>(
          m, e, n_rot, guard_hit);
    } else {
      small_sweep_from<N, T, EVECS, SCALED, P + 1, NZ, ONE>(m, e, n_rot, guard_hit);
    }
  }
}

// Sweep-start status of one lane: the skip test of every pair (in schedule order).
template <int N, typename T>
JPD_HD bool small_status(const T (&m)[SmallShape<N>::kTri]) {
  using Sh = SmallShape<N>;
  int key[Sh::kTri];
#pragma unroll
  for (int k = 0; k < Sh::kTri; ++k) key[k] = abs_key(m[k]);
  bool any_need = false;
#pragma unroll
  for (int P = 0; P < Sh::kSteps * Sh::kHalf; ++P) {
    const int A = Sh::player_a(P / Sh::kHalf, P % Sh::kHalf), B = Sh::player_b(P / Sh::kHalf, P % Sh::kHalf);
    if (A < N && B < N) {
      const int i = A < B ? A : B, j = A < B ? B : A;
      any_need = any_need | key_needs_rotation<T>(key[Sh::tri(i, j)], is_zero(m[Sh::tri(i, j)]),
                                                  key[Sh::tri(i, i)], key[Sh::tri(j, j)]);
    }
  }
  return any_need;
}

// Diagonalise the register-resident matrix. Returns the reference's return value:
// rotations + 1 (> 0) once a sweep starts with nothing left to rotate, 0 if that was not
// observed within max_num_sweeps sweeps and n > 1 (jacobi_pd.hpp:214-219).
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
  bool rescale = false;  // warp-uniform: a lane met an x outside the safe window
  for (int sweep = 0; sweep < max_num_sweeps; ++sweep) {
    const bool any_need = small_status<N, T>(m);
    converged = converged || !any_need;
    if (!JPD_WARP_ANY(any_need)) break;
    bool guard_hit = false;
#if defined(JPD_NO_FIRST_SWEEP_SPARSITY)
    if (rescale) small_sweep_from<N, T, EVECS, true, 0>(m, e, n_rot, guard_hit);
#else
    // sweep 0 starts from E = I (and `rescale` is still false): sparse eigenvector update
    if (sweep == 0) small_sweep_from<N, T, EVECS, false, 0, identity_mask<N>(), identity_mask<N>()>(m, e, n_rot, guard_hit);
    else if (rescale) small_sweep_from<N, T, EVECS, true, 0>(m, e, n_rot, guard_hit);
#endif
    else         small_sweep_from<N, T, EVECS, false, 0>(m, e, n_rot, guard_hit);
    rescale = rescale || JPD_WARP_ANY(guard_hit);
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

#ifndef JPD_SMALL_THREADS
#define JPD_SMALL_THREADS 128
#endif
#ifndef JPD_MINB_N3
#define JPD_MINB_N3 8
#endif
#ifndef JPD_MINB_N4
#define JPD_MINB_N4 6
#endif
template <int N, typename T> struct SmallLaunch {
  // resident blocks per SM the register allocator is asked to make room for
  static constexpr int kMinBlocks = (sizeof(T) == 8) ? (N <= 2 ? 8 : (N == 3 ? JPD_MINB_N3 : (N == 4 ? JPD_MINB_N4 : (N == 5 ? 4 : 2))))
                                                     : (N <= 4 ? 8 : (N == 5 ? 6 : 4));
};

// layout 0 (AoS): mat[b*N*N + i*N + j]; layout 1 (SoA-interleaved): mat[(i*N+j)*batch + b]
template <int N, typename T, bool SOA, bool EVECS>
__global__ void __launch_bounds__(JPD_SMALL_THREADS, SmallLaunch<N, T>::kMinBlocks)
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
  // The sort permutes DESTINATIONS, not data: source row r goes to output row pos[r].  (Moving
  // the rows through selects or a local-memory table costs ~200 more instructions per matrix;
  // the scattered 8-byte stores of one warp still fill whole sectors over the N rows and are
  // merged in L2 -- DRAM traffic stays at the algorithmic 8(N+N^2)+4 bytes, see profiles/.)
  int pos[N];
#pragma unroll
  for (int r = 0; r < N; ++r) {
    pos[r] = r;
    if (sorted) {
#pragma unroll
      for (int k = 0; k < N; ++k) pos[r] = (id[k] == r) ? k : pos[r];
    }
  }
#pragma unroll
  for (int r = 0; r < N; ++r) {
    if (SOA) evals[(long long)pos[r] * batch + b] = ev[r];
    else evals[b * N + pos[r]] = ev[r];
  }
  if (EVECS) {
#pragma unroll
    for (int r = 0; r < N; ++r) {
      T* row = SOA ? evecs + (long long)pos[r] * N * batch + b : evecs + b * (N * N) + pos[r] * N;
#pragma unroll
      for (int v = 0; v < N; ++v) {
        if (SOA) row[(long long)v * batch] = e[r * N + v];
        else row[v] = e[r * N + v];
      }
    }
  }
}

#endif  // __CUDACC__

}  // namespace jpd
