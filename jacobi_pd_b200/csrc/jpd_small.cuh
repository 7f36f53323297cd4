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
//  * Round 2 (instruction diet, DESIGN.md "tier 1"): the skip test of jacobi_pd.hpp:191 is
//    evaluated ONCE per sweep (small_status), not again per pair: inside a sweep every pair
//    whose off-diagonal entry is nonzero is rotated -- a pair the test would have skipped gets
//    a rotation that is the identity to working precision, which is what the reference's
//    `M[i][j] = 0` (:192) does to it as well.  A lane that has converged is frozen exactly
//    (its off-diagonal entries are zeroed), so results do not depend on warp neighbours.  The diagonal lives HALVED in registers
//    (exact), which removes two FP64 instructions from every rotation (see
//    small_rotate_pair).
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
  // added to every x (small_rotate_pair): 2^-1000 (fp64) / 2^-120 (fp32), a normal number that
  // is below half an ulp of every x inside the window
  JPD_HD static T tiny() { return pow2_from_field(sizeof(T) == 8 ? 23 : 7, T(0)); }
};

// the word that holds sign and exponent, as is (callers know x >= 0 or do not care)
JPD_HD int high_word(double x) {
#if defined(__CUDA_ARCH__)
  return __double2hiint(x);
#else
  std::uint64_t u; std::memcpy(&u, &x, 8);
  return (int)(u >> 32);
#endif
}
JPD_HD int high_word(float x) {
#if defined(__CUDA_ARCH__)
  return __float_as_int(x);
#else
  std::uint32_t u; std::memcpy(&u, &x, 4);
  return (int)u;
#endif
}

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
// The diagonal is passed HALVED (h = m/2), hence kMant+1; the key is clamped at 0 so that an
// exact zero off-diagonal (key 0) is never flagged, whatever the diagonal.  (Entries below
// 2^-1042 in magnitude -- high word 0 -- count as zero here.)
template <typename T>
JPD_HD int diag_skip_key(T half_diag) {
  const int lowered = abs_key(half_diag) - ((Fp<T>::kMantBits + 1) << KeyShift<T>::value);
  return lowered > 0 ? lowered : 0;
}

// (2w)^-1/2 for w in [1, 2] WITHOUT forming 2w: the seed comes from the high word with the
// exponent field bumped (one integer add), and the cubic correction of rsqrt_correct() is
// rewritten in eh = e/2 = 1/2 - w z^2:  z (1 + e/2 + 3/8 e^2) = z + (1 + 3/2 eh) (eh z).
JPD_HD double rsqrt_refined_of_double(double w) {
#if defined(__CUDA_ARCH__)
  double z;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(z) : "d"(__hiloint2double(__double2hiint(w) + 0x00100000, 0)));
  const double wz = w * z;
  const double eh = __fma_rn(-wz, z, 0.5);
  const double r2 = __fma_rn(1.5, eh, 1.0);
  const double ehz = eh * z;
  return __fma_rn(r2, ehz, z);
#else
  return 1.0 / std::sqrt(2.0 * w);
#endif
}
JPD_HD float rsqrt_refined_of_double(float w) {
#if defined(__CUDA_ARCH__)
  float z;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(z) : "f"(__int_as_float(__float_as_int(w) + 0x00800000)));
  const float eh = __fmaf_rn(-(w * z), z, 0.5f);
  return __fmaf_rn(z, eh, z);              // one Newton step: z (1 + e/2)
#else
  return (float)(1.0 / std::sqrt(2.0 * (double)w));
#endif
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

// One Jacobi rotation of the static pair (I,J), I<J, branch-free.
//
// m holds the diagonal HALVED: hI = mII/2, hJ = mJJ/2.  With D = hJ - hI = d/2:
//     x = a^2 + D^2 = (d^2 + 4a^2)/4,   P = x^-1/2 = 2/sqrt(d^2+4a^2)
//     w = 1 + |D| P = 1 + cos 2θ = 2 cos^2 θ   (one FMA; in [1,2])
//     U = (2w)^-1/2 = 1/(2 cos θ)              (2w: integer add on the exponent)
//     c = w U,   s = (a P) U sgn(d),   g = s U = tan θ / 2
//     hI' = hI - g a,  hJ' = hJ + g a          (jacobi_pd.hpp:337-338, halved)
// Same angle as CalcRot (jacobi_pd.hpp:233-256: |t| <= 1).  18 FP64-pipe instructions
// (2 refined rsqrt = 10, plus 8) + 2 for the diagonal; no division.
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
JPD_HD void small_rotate_pair(T (&m)[SmallShape<N>::kTri], T (&e)[EVECS ? N * N : 1], bool& out_of_range) {
  using Sh = SmallShape<N>;
  const T a = m[Sh::tri(I, J)];
  const T hi = m[Sh::tri(I, I)];
  const T hj = m[Sh::tri(J, J)];
  const bool a_nz = !is_zero(a);
  T dl = hj - hi;
  T as = a;
  // sign(t) = sign(a) sign(d).  For d == 0 the reference takes t = +1 whatever the sign of a
  // (jacobi_pd.hpp:237-239): kept for N == 2, where the rotation order IS the reference's, so
  // that even the unsorted output agrees; for N > 2 (d == +0) t = sign(a) -- either choice
  // zeroes the pivot, and the unsorted order is implementation defined there anyway.
  T sign_src = dl;
  if constexpr (N == 2) sign_src = is_zero(dl) ? a : dl;
  if (SCALED) rescale_for_rotation(dl, as);
  // + tiny: x > 0 even for a pair that is exactly diagonal already (a = D = 0), so that P
  // stays finite and s = g = 0 * P * U = 0 without a select.  tiny is far below the window
  // in which the un-rescaled variant accepts x, and below 2^-900 of the rescaled x in [1,8).
  const T x = jfma(as, as, jfma(dl, dl, SmallRange<T>::tiny()));
  // Un-rescaled variant: a nonzero pivot whose x left the window in which squares neither
  // overflow nor lose bits to underflow cannot be rotated here; small_solve() then discards
  // this pass for the whole warp and repeats it with the rescaling variant.
  if (!SCALED) {
    const bool in_window = (unsigned)(high_word(x) - (SmallRange<T>::kXMin << KeyShift<T>::value)) <
                           (unsigned)((SmallRange<T>::kXMax - SmallRange<T>::kXMin) << KeyShift<T>::value);
    out_of_range = out_of_range | (a_nz & !in_window);
  }
  const T p = rsqrt_refined(x);
  const T w = jfma(jabs(dl), p, T(1));
  const T u = rsqrt_refined_of_double(w);
  T c = w * u;
  const T s = xor_sign(as * p, sign_src) * u;
  const T g = s * u;
  // a == 0: the exact identity (s = g = 0 already; c would only be 1 to an ulp)
  c = a_nz ? c : T(1);
  // diagonal through t, pivot zeroed: jacobi_pd.hpp:337-338,347
  m[Sh::tri(I, I)] = jfma(-g, a, hi);
  m[Sh::tri(J, J)] = jfma(g, a, hj);
  m[Sh::tri(I, J)] = T(0);
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
JPD_HD void small_sweep_from(T (&m)[SmallShape<N>::kTri], T (&e)[EVECS ? N * N : 1], bool& out_of_range) {
  using Sh = SmallShape<N>;
  if constexpr (P < Sh::kSteps * Sh::kHalf) {
    constexpr int S = P / Sh::kHalf, K = P % Sh::kHalf;
    constexpr int A = Sh::player_a(S, K), B = Sh::player_b(S, K);
    if constexpr (A < N && B < N) {  // a pair with the dummy player (odd N) sits out
      constexpr int I = A < B ? A : B, J = A < B ? B : A;
      small_rotate_pair<N, T, EVECS, SCALED, I, J, NZ, ONE>(m, e, out_of_range);
      small_sweep_from<N, T, EVECS, SCALED, P + 1, mask_after_nz<N>(NZ, I, J), mask_after_one<N>(ONE, I, J)>(
          m, e, out_of_range);
    } else {
      small_sweep_from<N, T, EVECS, SCALED, P + 1, NZ, ONE>(m, e, out_of_range);
    }
  }
}

// Sweep-start status of one lane: the skip test of every pair (m holds the halved diagonal).
template <int N, typename T>
JPD_HD bool small_status(const T (&m)[SmallShape<N>::kTri]) {
  using Sh = SmallShape<N>;
  int kd[N];
#pragma unroll
  for (int i = 0; i < N; ++i) kd[i] = diag_skip_key<T>(m[Sh::tri(i, i)]);
  bool any_need = false;
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = i + 1; j < N; ++j) {
      const int ka = abs_key(m[Sh::tri(i, j)]);
      any_need = any_need | (ka > kd[i]) | (ka > kd[j]);   // ka > min(kd[i], kd[j])
    }
  return any_need;
}

// One pass over the register-resident matrix.  Returns the reference's return value:
// rotations + 1 (> 0) once a sweep starts with nothing left to rotate, 0 if that was not
// observed within max_num_sweeps sweeps and n > 1 (jacobi_pd.hpp:214-219).
// SCALED = false is the common pass; a lane that meets a pivot it cannot rotate without
// rescaling sets out_of_range (its own result is then garbage) and no longer counts in the
// warp's vote to go on sweeping.
template <int N, typename T, bool EVECS, bool SCALED>
JPD_HD int small_solve_pass(T (&m)[SmallShape<N>::kTri], T (&e)[EVECS ? N * N : 1], int max_num_sweeps,
                            bool& out_of_range) {
  if (EVECS) {
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) e[i * N + j] = (i == j) ? T(1) : T(0);  // :172-178
  }
  out_of_range = false;
  if (N == 1) return 1;
#pragma unroll
  for (int i = 0; i < N; ++i) m[SmallShape<N>::tri(i, i)] *= T(0.5);  // exact (see small_rotate_pair)
  constexpr int kPairs = N * (N - 1) / 2;
  int n_rot = 0;
  bool converged = false;
  for (int sweep = 0; sweep < max_num_sweeps; ++sweep) {
    const bool any_need = small_status<N, T>(m);
    converged = converged || !any_need;
    // a lane that already gave up (out_of_range) does not keep its warp sweeping
    if (!JPD_WARP_ANY(any_need && !out_of_range)) break;
    // A lane whose matrix has converged but whose warp goes on is FROZEN: its negligible
    // off-diagonal entries become exact zeros (what the reference does to its last pivot,
    // jacobi_pd.hpp:192), so every rotation of the coming sweeps is the exact identity for it
    // (a == 0: c = 1, s = g = 0).  A matrix's result therefore never depends on the matrices
    // that happen to share its warp -- bit for bit, whatever the batch or chunk size.
    if (!any_need) {
#pragma unroll
      for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = i + 1; j < N; ++j) m[SmallShape<N>::tri(i, j)] = T(0);
    }
    // rotations attempted: every pair of a sweep this lane still needed (for n = 2 exactly the
    // reference's count; the n_iters contract is "> 0 iff converged", jacobi_pd.hpp:73-74)
    n_rot += converged ? 0 : kPairs;
    if (SCALED) {
      small_sweep_from<N, T, EVECS, true, 0>(m, e, out_of_range);
    } else {
      // sweep 0 starts from E = I: sparse eigenvector update
      if (sweep == 0) small_sweep_from<N, T, EVECS, false, 0, identity_mask<N>(), identity_mask<N>()>(m, e, out_of_range);
      else            small_sweep_from<N, T, EVECS, false, 0>(m, e, out_of_range);
    }
  }
#pragma unroll
  for (int i = 0; i < N; ++i) m[SmallShape<N>::tri(i, i)] *= T(2);
  return converged ? n_rot + 1 : 0;
}

// Diagonalise one register-resident matrix per lane: `load(m)` fills the upper triangle,
// `emit(m, e, n_iters)` consumes the result (m: eigenvalues on the diagonal, e: eigenvector
// rows).  If a lane of the warp met a pivot that needs rescaling (huge, tiny or denormal
// entries -- rare), the lanes that did NOT emit their finished result at once and only the
// others reload and repeat with the rescaling variant: a matrix never changes the bits of its
// neighbours' results, and no state has to be kept in registers for the rare case.
template <int N, typename T, bool EVECS, typename Load, typename Emit>
JPD_HD void small_solve(int max_num_sweeps, Load load, Emit emit) {
  T m[SmallShape<N>::kTri];
  T e[EVECS ? N * N : 1];
  load(m);
  bool out_of_range;
  int iters = small_solve_pass<N, T, EVECS, false>(m, e, max_num_sweeps, out_of_range);
  if (JPD_WARP_ANY(out_of_range)) {
    if (!out_of_range) { emit(m, e, iters); return; }
    bool unused;
    load(m);
    iters = small_solve_pass<N, T, EVECS, true>(m, e, max_num_sweeps, unused);
  }
  emit(m, e, iters);
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

// Sort key: the reference's four orderings (jacobi_pd.hpp:483-499) all become "larger key
// first" after clearing and/or flipping the sign bit -- integer pipe, one LOP3 per key.
JPD_HD double sort_key(double v, int clear_mask, int flip_mask) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double((__double2hiint(v) & clear_mask) ^ flip_mask, __double2loint(v));
#else
  std::uint64_t u; std::memcpy(&u, &v, 8);
  u = (u & (((std::uint64_t)(std::uint32_t)clear_mask << 32) | 0xffffffffull)) ^ ((std::uint64_t)(std::uint32_t)flip_mask << 32);
  double r; std::memcpy(&r, &u, 8); return r;
#endif
}
JPD_HD float sort_key(float v, int clear_mask, int flip_mask) {
#if defined(__CUDA_ARCH__)
  return __int_as_float((__float_as_int(v) & clear_mask) ^ flip_mask);
#else
  std::uint32_t u; std::memcpy(&u, &v, 4);
  u = (u & (std::uint32_t)clear_mask) ^ (std::uint32_t)flip_mask;
  float r; std::memcpy(&r, &u, 4); return r;
#endif
}

// pos[r] = output position of source slot r under SortRows (jacobi_pd.hpp:477-508).
// Without ties among the keys the selection sort's result is simply the ranking, which costs
// two instructions per pair; a lane that does hold equal keys (exactly repeated eigenvalues,
// +-lambda under the ABS criteria) replays the reference's selection sort, whose tie order is
// not the stable one.  Returns false (pos = identity) for DO_NOT_SORT / unknown criteria.
template <int N, typename T>
JPD_HD bool small_sort_positions(const T (&ev)[N], int sort_criteria, int (&pos)[N]) {
#pragma unroll
  for (int r = 0; r < N; ++r) pos[r] = r;
  if (sort_criteria < SORT_DECREASING_EVALS || sort_criteria > SORT_INCREASING_ABS_EVALS)
    return false;  // no-op (jacobi_pd.hpp:500-501)
  const int clear_mask = sort_criteria >= SORT_DECREASING_ABS_EVALS ? 0x7fffffff : -1;
  const int flip_mask = (sort_criteria == SORT_INCREASING_EVALS || sort_criteria == SORT_INCREASING_ABS_EVALS)
                            ? (int)0x80000000 : 0;
  T key[N];
#pragma unroll
  for (int k = 0; k < N; ++k) { key[k] = sort_key(ev[k], clear_mask, flip_mask); pos[k] = 0; }
  bool tie = false;
#pragma unroll
  for (int k = 0; k < N; ++k)
#pragma unroll
    for (int r = k + 1; r < N; ++r) {
      const bool first = key[k] >= key[r];
      tie = tie | (key[k] == key[r]);
      pos[r] += first ? 1 : 0;
      pos[k] += first ? 0 : 1;
    }
  if (tie) {
    int id[N];
    small_sort_ids<N, T>(ev, sort_criteria, id);
#pragma unroll
    for (int r = 0; r < N; ++r)
#pragma unroll
      for (int k = 0; k < N; ++k) pos[r] = (id[k] == r) ? k : pos[r];
  }
  return true;
}

#if defined(__CUDACC__)

#ifndef JPD_SMALL_THREADS
#define JPD_SMALL_THREADS 128
#endif
#ifndef JPD_MINB_N3
#define JPD_MINB_N3 10
#endif
#ifndef JPD_MINB_N4
#define JPD_MINB_N4 6
#endif
template <int N, typename T> struct SmallLaunch {
  // resident blocks per SM the register allocator is asked to make room for
  static constexpr int kMinBlocks = (sizeof(T) == 8) ? (N <= 2 ? 8 : (N == 3 ? JPD_MINB_N3 : (N == 4 ? JPD_MINB_N4 : (N == 5 ? 4 : 2))))
                                                     : (N <= 4 ? 8 : (N == 5 ? 6 : 4));
};

template <int N, typename T, bool SOA>
__device__ __forceinline__ const T* small_entry_ptr(const T* __restrict__ mat, long long batch, long long b, int i, int j) {
  return SOA ? mat + (long long)(i * N + j) * batch + b : mat + b * (N * N) + (i * N + j);
}

// Epilogue of one matrix: eigenvalues from the diagonal (:207-209), SortRows (:477-508) as a
// permutation of DESTINATIONS, not data: source row r goes to output row pos[r].  (Moving the
// rows through selects or a local-memory table costs ~200 more instructions per matrix; the
// scattered 8-byte stores of one warp still fill whole sectors over the N rows and are merged
// in L2 -- DRAM traffic stays at the algorithmic 8(N+N^2)+4 bytes, see profiles/.)
template <int N, typename T, bool SOA, bool EVECS>
__device__ __forceinline__ void small_emit(const T (&m)[SmallShape<N>::kTri], const T (&e)[EVECS ? N * N : 1], int iters,
                                           bool live, long long b, long long batch, int sort_criteria,
                                           T* __restrict__ evals, T* __restrict__ evecs, int* __restrict__ n_iters) {
  using Sh = SmallShape<N>;
  T ev[N];
#pragma unroll
  for (int i = 0; i < N; ++i) ev[i] = m[Sh::tri(i, i)];
  int pos[N];
  small_sort_positions<N, T>(ev, sort_criteria, pos);
  if (!live) return;
  if (n_iters) n_iters[b] = iters;
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

// layout 0 (AoS): mat[b*N*N + i*N + j]; layout 1 (SoA-interleaved): mat[(i*N+j)*batch + b]
// One matrix per thread.  (A persistent variant -- warps drawing tiles from a ticket counter,
// the next tile's entries prefetched into shared memory with cp.async -- was built and measured
// in round 2: it keeps 5.9 instead of 5.5 warps per scheduler resident and removes the
// load-latency stalls, and is not faster, because the kernel is bound by the issue port: an
// FP64 instruction occupies it for two cycles, see DESIGN.md "tier 1".)
template <int N, typename T, bool SOA, bool EVECS>
__global__ void __launch_bounds__(JPD_SMALL_THREADS, SmallLaunch<N, T>::kMinBlocks)
jpd_small_kernel(const T* __restrict__ mat, T* __restrict__ evals, T* __restrict__ evecs,
                 int* __restrict__ n_iters, long long batch, int sort_criteria,
                 int max_num_sweeps) {
  using Sh = SmallShape<N>;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = gid < batch;
  const long long b = live ? gid : batch - 1;  // tail lanes shadow the last matrix (no stores)
  auto load = [&](T (&m)[Sh::kTri]) {
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = i; j < N; ++j) m[Sh::tri(i, j)] = __ldg(small_entry_ptr<N, T, SOA>(mat, batch, b, i, j));
  };
  auto emit = [&](const T (&m)[Sh::kTri], const T (&e)[EVECS ? N * N : 1], int iters) {
    small_emit<N, T, SOA, EVECS>(m, e, iters, live, b, batch, sort_criteria, evals, evecs, n_iters);
  };
  small_solve<N, T, EVECS>(max_num_sweeps, load, emit);
}

#endif  // __CUDACC__

}  // namespace jpd
