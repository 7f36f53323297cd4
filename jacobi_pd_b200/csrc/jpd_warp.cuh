// jpd_warp.cuh -- tier 2: one warp per matrix (17 <= n <= 32), two matrices per warp
// (9 <= n <= 16) or four (7 <= n <= 8).
//
// Path being replaced: Jacobi::Diagonalize, /root/reference/include/jacobi_pd.hpp:159-220
// (pivot search :427-468, CalcRot :233-256, ApplyRot :327-393, ApplyRotLeft :409-419,
//  SortRows :477-508).
//
// Design ("tier 2"):
//  * Odd-even (Brent-Luk linear-array) parallel ordering on POSITIONS.  A sweep is m steps
//    (m = n rounded up to even) alternating two static pairings,
//        pattern A: positions (0,1)(2,3)...      pattern B: positions (1,2)(3,4)...
//    and every rotation also exchanges its two positions, so that after m steps every pair
//    of indices has met exactly once (and the order of the indices is reversed).  The
//    exchange costs nothing: the rotation writes (s x + c y, c x - s y) instead of
//    (c x - s y, s x + c y).  Nothing is ever addressed by a data-dependent index.
//  * Because the pairings are static, the eigenvector accumulation lives in REGISTERS:
//    lane v owns column v of E, its rows are registers with compile-time names, and a
//    step is NP/2 independent 2x2 row operations per lane (4 FP64 instructions each).
//  * M lives in shared memory, upper triangle only, with one phantom row/column on either
//    side (-1 and m): pattern B is then pattern A shifted by one position, its two
//    unpaired end positions being paired with a phantom through the constant "rotation"
//    (c,s) = (0,-+1), which reproduces them unchanged.  Columns are stored even/odd split
//    so that the lanes of a warp (consecutive 2x2 blocks of one block row) always touch
//    consecutive addresses: no bank conflicts in either pattern.
//  * A step = phase 1 (one lane per pair: the reference's skip test :191, rotation, the
//    pair's own three entries, (c,s) to shared memory) + phase 2 (all lanes: each 2x2
//    block {pair K} x {pair L}, K<L, gets G_K B G_L^T once: 16 FP64 instructions; E rows).
//  * Convergence = the reference's test applied to ALL pairs at the start of each sweep
//    (cheap next to a sweep), so no empty sweep is ever executed.
//  * Return value as the reference: rotations + 1, or 0 when max_num_sweeps sweeps were
//    not enough (:214-219).
#pragma once
#include "jpd_common.cuh"

namespace jpd {

template <int NP> struct WarpShape {
  static constexpr int kH = NP / 2 + 1;            // column slots per parity class
  static constexpr int kLd = 2 * kH;               // NP + 2 elements per stored row
  static constexpr int kRows = NP + 2;             // positions -1 .. NP
  static constexpr int kMElems = kRows * kLd;
  static constexpr int kPairsMax = NP / 2 + 1;
  static constexpr int kGroups = 32 / NP;          // matrices per warp
  static constexpr int kBlkMaxA = (NP / 2) * (NP / 2 - 1) / 2;
  static constexpr int kBlkMaxB = (NP / 2 + 1) * (NP / 2) / 2;
  static constexpr int kItA = (kBlkMaxA + NP - 1) / NP;
  static constexpr int kItB = (kBlkMaxB + NP - 1) / NP;
  // per-group shared memory, in elements of T: M, (c,s) pairs, eigenvalues; then ints
  static constexpr int kGroupElems = kMElems + 2 * kPairsMax + NP + 2;
  JPD_CX static constexpr size_t group_bytes(size_t elem) {
    return ((kGroupElems * elem + NP * sizeof(int)) + 15) / 16 * 16;
  }
  JPD_CX static constexpr size_t table_bytes() {
    return ((size_t)(kBlkMaxA + kBlkMaxB) * sizeof(unsigned) + 15) / 16 * 16;
  }
  JPD_CX static constexpr size_t block_bytes(size_t elem, int warps) {
    return table_bytes() + (size_t)warps * kGroups * group_bytes(elem);
  }
  // storage offset of element (r,c), r,c in [-1, NP]
  JPD_HD static int addr(int r, int c) {
    const int x = c + 1;
    return (r + 1) * kLd + (x >> 1) + (x & 1) * kH;
  }
};

#if defined(__CUDACC__)

template <typename T> struct Pair2;
template <> struct Pair2<double> { using type = double2; };
template <> struct Pair2<float> { using type = float2; };

// order-preserving integer image of a floating-point key (total order, NaNs at the ends)
__device__ __forceinline__ long long sortable_bits(double x) {
  const long long k = __double_as_longlong(x);
  return k ^ ((k >> 63) & 0x7fffffffffffffffLL);
}
__device__ __forceinline__ long long sortable_bits(float x) {
  const int k = __float_as_int(x);
  return (long long)(k ^ ((k >> 31) & 0x7fffffff));
}

// flat index over {(K,L): 0 <= K < L < np}, row-major in K  ->  (K,L)
__device__ __forceinline__ void tri_decode(int f, int np, int& K, int& L) {
  int k = 0, rem = f;
  while (rem >= np - 1 - k) { rem -= np - 1 - k; ++k; }
  K = k;
  L = k + 1 + rem;
}

// One step of the odd-even ordering.  SH = 0: pattern A, pairs (2K, 2K+1), K < m/2.
// SH = -1: pattern B, pairs (2K-1, 2K), K <= m/2, pairs 0 and m/2 half phantom.
template <typename T, int NP, bool EVECS, int SH>
__device__ __forceinline__ void warp_step(T* __restrict__ M, T* __restrict__ cs,
                                          const unsigned* __restrict__ tbl, T (&e)[EVECS ? NP : 1],
                                          int m, int sub, int& n_rot) {
  using Sh = WarpShape<NP>;
  using T2 = typename Pair2<T>::type;
  constexpr int kIt = (SH == 0) ? Sh::kItA : Sh::kItB;
  const int np = (m >> 1) + (SH == 0 ? 0 : 1);
  const int nblk = np * (np - 1) / 2;
  T2* cs2 = reinterpret_cast<T2*>(cs);

  // ---- phase 1: one lane per pair -- skip test (:191), CalcRot (:233-256), the pair's
  // own 2x2 block (:337-338,347) with the two positions exchanged
  if (sub < np) {
    const int p = 2 * sub + SH, q = p + 1;
    T c = T(1), s = T(0);
    if (SH != 0 && p < 0) { c = T(0); s = T(-1); }        // (phantom, 0): keeps position 0
    else if (SH != 0 && q >= m) { c = T(0); s = T(1); }   // (m-1, phantom): keeps position m-1
    else {
      T* pp = M + Sh::addr(p, p);
      T* qq = M + Sh::addr(q, q);
      T* pq = M + Sh::addr(p, q);
      const T mpp = *pp, mqq = *qq, a = *pq;
      T npp = mqq, nqq = mpp;
      if (pair_needs_rotation(a, mpp, mqq)) {
        const T d = mqq - mpp;
        T t;
        jacobi_rotation(d, a, rotation_range_unsafe(d, a), c, s, t);
        const T g = t * a;
        npp = mqq + g;
        nqq = mpp - g;
        *pq = T(0);
        ++n_rot;
      }
      *pp = npp;
      *qq = nqq;
    }
    T2 v; v.x = c; v.y = s;
    cs2[sub] = v;
  }
  __syncwarp();

  // ---- phase 2a: 2x2 blocks rows {pair K} x cols {pair L}, K < L:  B <- G_K B G_L^T with
  // G = [[s, c], [c, -s]] (ApplyRot :350-390 for both rotations, positions exchanged)
#pragma unroll
  for (int it = 0; it < kIt; ++it) {
    const int f = it * NP + sub;
    if (f < nblk) {
      const unsigned ent = tbl[f];
      T* b = M + (ent & 0xffffu);
      const T2 rk = cs2[(ent >> 16) & 0xffu];
      const T2 rl = cs2[ent >> 24];
      constexpr int o00 = (SH == 0) ? Sh::kH : 0;
      constexpr int o01 = (SH == 0) ? 1 : Sh::kH;
      const T b00 = b[o00], b01 = b[o01], b10 = b[Sh::kLd + o00], b11 = b[Sh::kLd + o01];
      const T ck = rk.x, sk = rk.y, cl = rl.x, sl = rl.y;
      const T u00 = jfma(sk, b00, ck * b10), u01 = jfma(sk, b01, ck * b11);
      const T u10 = jfma(ck, b00, -(sk * b10)), u11 = jfma(ck, b01, -(sk * b11));
      b[o00] = jfma(sl, u00, cl * u01);
      b[o01] = jfma(cl, u00, -(sl * u01));
      b[Sh::kLd + o00] = jfma(sl, u10, cl * u11);
      b[Sh::kLd + o01] = jfma(cl, u10, -(sl * u11));
    }
  }
  // ---- phase 2b: eigenvector rows (ApplyRotLeft :414-418), registers only
  if (EVECS) {
#pragma unroll
    for (int K = (SH == 0 ? 0 : 1); K < NP / 2; ++K) {
      const int p = 2 * K + SH, q = p + 1;   // compile-time register names
      if (K < np + SH) {                      // A: K < m/2   B: K <= m/2 - 1 (real pairs only)
        const T2 r = cs2[K];
        const T x = e[p], y = e[q];
        e[p] = jfma(r.y, x, r.x * y);
        e[q] = jfma(r.x, x, -(r.y * y));
      }
    }
  }
  __syncwarp();
}

#ifndef JPD_WARP_THREADS
#define JPD_WARP_THREADS 128
#endif

template <typename T, int NP, bool EVECS>
__global__ void __launch_bounds__(JPD_WARP_THREADS)
jpd_warp_kernel(const T* __restrict__ mat, T* __restrict__ evals, T* __restrict__ evecs,
                int* __restrict__ n_iters, int n, long long batch, int soa, int sort_criteria,
                int max_num_sweeps) {
  using Sh = WarpShape<NP>;
  constexpr int G = Sh::kGroups;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int grp = lane / NP, sub = lane % NP;
  const unsigned grp_mask = (NP == 32) ? 0xffffffffu : (((1u << NP) - 1u) << (grp * NP));
  const int m = n + (n & 1);
  const int npA = m >> 1, npB = npA + 1;

  // ---- block-wide tables: flat block index -> (offset of the block, K, L), both patterns
  unsigned* tblA = reinterpret_cast<unsigned*>(smem_raw);
  unsigned* tblB = tblA + Sh::kBlkMaxA;
  for (int f = threadIdx.x; f < npA * (npA - 1) / 2; f += blockDim.x) {
    int K, L; tri_decode(f, npA, K, L);
    tblA[f] = (unsigned)((2 * K + 1) * Sh::kLd + L) | ((unsigned)K << 16) | ((unsigned)L << 24);
  }
  for (int f = threadIdx.x; f < npB * (npB - 1) / 2; f += blockDim.x) {
    int K, L; tri_decode(f, npB, K, L);
    tblB[f] = (unsigned)((2 * K) * Sh::kLd + L) | ((unsigned)K << 16) | ((unsigned)L << 24);
  }
  __syncthreads();

  unsigned char* gbase = smem_raw + Sh::table_bytes() + (size_t)(warp * G + grp) * Sh::group_bytes(sizeof(T));
  T* M = reinterpret_cast<T*>(gbase);
  T* cs = M + Sh::kMElems;
  T* ev = cs + 2 * Sh::kPairsMax;
  int* perm = reinterpret_cast<int*>(ev + NP + 2);

  const long long b = ((long long)blockIdx.x * nwarps + warp) * G + grp;
  const bool live = b < batch;

  // ---- load: upper triangle only (jacobi_pd.hpp:167-171); everything else zero
  for (int i = sub; i < Sh::kMElems; i += NP) M[i] = T(0);
  __syncwarp();
  if (live) {
    for (int r = 0; r < n; ++r) {
      if (sub >= r && sub < n) {
        const T v = soa ? mat[((size_t)r * n + sub) * (size_t)batch + (size_t)b]
                        : mat[(size_t)b * n * n + (size_t)r * n + sub];
        M[Sh::addr(r, sub)] = v;
      }
    }
  }
  T e[EVECS ? NP : 1];
  if (EVECS) {
#pragma unroll
    for (int k = 0; k < NP; ++k) e[k] = (k == sub) ? T(1) : T(0);   // :172-178
  }
  __syncwarp();

  // ---- sweeps
  int n_rot = 0, sweeps = 0;
  bool need = false;
  for (;;) {
    // the reference's convergence test (:191) on every pair (i, i+d mod m)
    need = false;
    if (sub < m) {
      const T dii = M[Sh::addr(sub, sub)];
      for (int d = 1; d <= (m >> 1); ++d) {
        int j = sub + d; if (j >= m) j -= m;
        const int lo = sub < j ? sub : j, hi = sub < j ? j : sub;
        need = need || pair_needs_rotation(M[Sh::addr(lo, hi)], dii, M[Sh::addr(j, j)]);
      }
    }
    const unsigned vote = __ballot_sync(0xffffffffu, need);
    need = (vote & grp_mask) != 0u;
    if (vote == 0u || sweeps >= max_num_sweeps) break;
    for (int step = 0; step < m; step += 2) {
      warp_step<T, NP, EVECS, 0>(M, cs, tblA, e, m, sub, n_rot);
      warp_step<T, NP, EVECS, -1>(M, cs, tblB, e, m, sub, n_rot);
    }
    ++sweeps;
  }
  const bool converged = !need || n == 1;

  // ---- rotations of this matrix: sum over the lanes of its group
#pragma unroll
  for (int off = NP / 2; off > 0; off >>= 1) n_rot += __shfl_xor_sync(0xffffffffu, n_rot, off);

  // ---- eigenvalues (:207-209).  A sweep reverses the order of the positions, so for odd n
  // the padded index sits at position 0 after an odd number of sweeps, else at m - 1.
  const int first = ((n & 1) && (sweeps & 1)) ? 1 : 0;
  T my_ev = T(0);
  if (sub < n) { my_ev = M[Sh::addr(first + sub, first + sub)]; ev[sub] = my_ev; }
  __syncwarp();
  // sort: output position of slot `sub` = number of slots that precede it.  Keys are
  // compared as order-preserving integers (ties by slot), so the ranks are a permutation
  // whatever the values; the order of the KEYS is the one SortRows (:477-508) produces.
  if (sub < n) {
    int rank = sub;
    if (sort_criteria >= SORT_DECREASING_EVALS && sort_criteria <= SORT_INCREASING_ABS_EVALS) {
      const bool use_abs = sort_criteria >= SORT_DECREASING_ABS_EVALS;
      const bool decreasing = (sort_criteria == SORT_DECREASING_EVALS) || (sort_criteria == SORT_DECREASING_ABS_EVALS);
      const long long mine = sortable_bits(use_abs ? jabs(my_ev) : my_ev);
      rank = 0;
      for (int j = 0; j < n; ++j) {
        const T o = ev[j];
        const long long other = sortable_bits(use_abs ? jabs(o) : o);
        const bool before = decreasing ? (other > mine) : (other < mine);
        rank += (before || (other == mine && j < sub)) ? 1 : 0;
      }
    }
    perm[rank] = sub;
  }
  __syncwarp();

  // ---- store (all reads of M are done: its storage now stages the eigenvector rows)
  if (live && sub < n) {
    const T v = ev[perm[sub]];
    if (soa) evals[(size_t)sub * (size_t)batch + (size_t)b] = v;
    else evals[(size_t)b * n + sub] = v;
  }
  if (live && sub == 0 && n_iters) n_iters[b] = converged ? n_rot + 1 : 0;
  if (EVECS) {
    T* Es = M;
#pragma unroll
    for (int k = 0; k < NP; ++k) Es[k * NP + sub] = e[k];
    __syncwarp();
    if (live && sub < n) {
      for (int k = 0; k < n; ++k) {
        const T v = Es[(first + perm[k]) * NP + sub];
        if (soa) evecs[((size_t)k * n + sub) * (size_t)batch + (size_t)b] = v;
        else evecs[(size_t)b * n * n + (size_t)k * n + sub] = v;
      }
    }
  }
}

#endif  // __CUDACC__

}  // namespace jpd
