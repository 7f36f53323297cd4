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
//  * M lives in shared memory, upper triangle only, columns stored even/odd split.  A step
//    = phase 1 (one lane per pair: the reference's skip test :191, rotation, the pair's own
//    three entries, (c,s) to shared memory) + phase 2 (all lanes: each 2x2 block
//    {pair K} x {pair L}, K<L, gets G_K B G_L^T once: 16 FP64 instructions; in pattern B
//    also the 1x2 / 2x1 strips of the two unpaired end positions; E rows).
//  * Shared-memory banks: with a row pitch == 3 (mod 8) elements the bank of a block is
//    (6K + L) mod 16; the lane that handles a block is chosen as exactly that residue and
//    blocks of one residue class go to different half-warp slots, so every access of every
//    half-warp touches 16 different banks (tables built once per thread block).
//  * Convergence = the reference's test applied to ALL pairs at the start of each sweep
//    (cheap next to a sweep), so no empty sweep is ever executed.
//  * Return value as the reference: rotations + 1, or 0 when max_num_sweeps sweeps were
//    not enough (:214-219).
#pragma once
#include "jpd_common.cuh"

namespace jpd {

template <int NP> struct WarpShape {
  static constexpr int kH = NP / 2;                               // column slots per parity class
  static constexpr int kLd = (NP == 8) ? 9 : NP + 3;              // row pitch: 35, 19 (== 3 mod 8), 9
  static constexpr int kBankMul = (2 * kLd) & 15;                 // bank(block K,L) = kBankMul*K + L
  static constexpr int kMElems = NP * kLd;
  static constexpr int kPairsMax = NP / 2;
  static constexpr int kGroups = 32 / NP;                         // matrices per warp
  static constexpr int kClasses = NP < 16 ? NP : 16;              // bank residue classes = lanes of a half warp
  static constexpr int kHalves = NP / kClasses;                   // half warps per matrix
  static constexpr int kIt = (NP == 32) ? 4 : (NP == 16 ? 2 : 1); // block iterations per step (both patterns)
  static constexpr int kTblEntries = kIt * NP;
  // per-group shared memory, in elements of T: M | (c,s) interleaved | c | s | eigenvalues
  static constexpr int kGroupElems = kMElems + 2 * kPairsMax + kPairsMax + kPairsMax + NP;
  JPD_CX static constexpr size_t group_bytes(size_t elem) {
    return ((kGroupElems * elem + NP * sizeof(int)) + 15) / 16 * 16;
  }
  JPD_CX static constexpr size_t table_bytes() { return (size_t)2 * kTblEntries * sizeof(unsigned); }
  JPD_CX static constexpr size_t block_bytes(size_t elem, int warps) {
    return table_bytes() + (size_t)warps * kGroups * group_bytes(elem);
  }
  // storage offset of element (r,c), 0 <= r,c < NP
  JPD_HD static int addr(int r, int c) { return r * kLd + (c >> 1) + (c & 1) * kH; }
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

// Schedule table of one pattern: entry [it*NP + lane] = offset of the block's first element
// | K << 16 | L << 24, or 0xffffffff when the lane idles in that iteration.  `off` is the
// position of pair 0's first index (0: pattern A, 1: pattern B), `npr` the number of pairs.
template <int NP>
__device__ __forceinline__ void build_schedule(unsigned* tbl, int npr, int off) {
  using Sh = WarpShape<NP>;
  for (int i = threadIdx.x; i < Sh::kTblEntries; i += blockDim.x) tbl[i] = 0xffffffffu;
  __syncthreads();
  if (threadIdx.x < Sh::kClasses) {
    const int cl = threadIdx.x;
    int slot = 0;
    for (int K = 0; K < npr; ++K) {
      const int L = (cl - Sh::kBankMul * K) & (Sh::kClasses - 1);
      if (L > K && L < npr) {
        const int idx = (slot / Sh::kHalves) * NP + (slot % Sh::kHalves) * Sh::kClasses + cl;
        if (idx < Sh::kTblEntries)
          tbl[idx] = (unsigned)Sh::addr(2 * K + off, 2 * L + off) | ((unsigned)K << 16) | ((unsigned)L << 24);
        ++slot;
      }
    }
  }
}

// One step of the odd-even ordering.  OFF = 0: pattern A, pairs (2j, 2j+1), j < m/2.
// OFF = 1: pattern B, pairs (2j+1, 2j+2), j < m/2 - 1; positions 0 and m-1 sit out.
template <typename T, int NP, bool EVECS, int OFF, bool FULL>
__device__ __forceinline__ void warp_step(T* __restrict__ M, T* __restrict__ cs, T* __restrict__ cc,
                                          T* __restrict__ ss, const unsigned* __restrict__ tbl,
                                          T (&e)[EVECS ? NP : 1], int m, int sub, int& n_rot) {
  using Sh = WarpShape<NP>;
  using T2 = typename Pair2<T>::type;
  const int npr = FULL ? NP / 2 - OFF : (m >> 1) - OFF;   // FULL: m == NP, no runtime guards
  T2* cs2 = reinterpret_cast<T2*>(cs);
  // offsets of the block's four elements from its first one (columns are even/odd split)
  constexpr int o01 = (OFF == 0) ? Sh::kH : 1 - Sh::kH;
  constexpr int o10 = Sh::kLd;
  constexpr int o11 = Sh::kLd + o01;

  // ---- phase 1: one lane per pair -- skip test (:191), CalcRot (:233-256), the pair's
  // own 2x2 block (:337-338,347) with the two positions exchanged
  if (sub < npr) {
    const int p = 2 * sub + OFF;
    T* pp = M + Sh::addr(p, p);
    T* pq = pp + o01;                  // (p, p+1)
    T* qq = pp + o11;                  // (p+1, p+1)
    const T mpp = *pp, mqq = *qq, a = *pq;
    T c, s, npp, nqq;
    if (pivot_exchange_rotation(mpp, mqq, a, c, s, npp, nqq)) {
      *pq = T(0);
      ++n_rot;
    }
    *pp = npp;
    *qq = nqq;
    T2 v; v.x = c; v.y = s;
    cs2[sub] = v;
    cc[sub] = c;
    ss[sub] = s;
  }
  __syncwarp();

  // ---- phase 2a: 2x2 blocks rows {pair K} x cols {pair L}, K < L:  B <- G_K B G_L^T with
  // G = [[s, c], [c, -s]] (ApplyRot :350-390 for both rotations, positions exchanged)
#pragma unroll
  for (int it = 0; it < Sh::kIt; ++it) {
    const unsigned ent = tbl[it * NP + sub];
    if (ent != 0xffffffffu) {
      T* b = M + (ent & 0xffffu);
      const int K = (ent >> 16) & 0xffu, L = ent >> 24;
      const T ck = cc[K], sk = ss[K], cl = cc[L], sl = ss[L];
      T b00 = b[0], b01 = b[o01], b10 = b[o10], b11 = b[o11];
      exchange_rotate_block(b00, b01, b10, b11, ck, sk, cl, sl);
      b[0] = b00; b[o01] = b01; b[o10] = b10; b[o11] = b11;
    }
  }
  // ---- pattern B: the unpaired end positions.  Row 0 meets the column pair j (first half
  // of the lanes), column m-1 meets the row pair j (second half); same 2-vector update.
  if (OFF == 1) {
    const int j = sub < NP / 2 ? sub : sub - NP / 2;
    if (j < npr) {
      T* x0 = (sub < NP / 2) ? M + Sh::addr(0, 2 * j + 1) : M + Sh::addr(2 * j + 1, (FULL ? NP : m) - 1);
      T* x1 = (sub < NP / 2) ? x0 + o01 : x0 + o10;
      T v0 = *x0, v1 = *x1;
      exchange_rotate_pair(v0, v1, cc[j], ss[j]);
      *x0 = v0; *x1 = v1;
    }
  }
  // ---- phase 2b: eigenvector rows (ApplyRotLeft :414-418), registers only
  if (EVECS) {
#pragma unroll
    for (int j = 0; j < NP / 2 - OFF; ++j) {
      const int p = 2 * j + OFF, q = p + 1;   // compile-time register names
      if (FULL || j < npr) {
        const T2 r = cs2[j];
        exchange_rotate_pair(e[p], e[q], r.x, r.y);
      }
    }
  }
  __syncwarp();
}

#ifndef JPD_WARP_THREADS
#define JPD_WARP_THREADS 128
#endif
// resident blocks per SM the register allocator is asked to make room for
template <typename T, int NP> struct WarpLaunch {
  static constexpr int kMinBlocks = sizeof(T) == 8 ? (NP == 32 ? 4 : (NP == 16 ? 7 : 8))
                                                   : (NP == 32 ? 6 : 8);
};

// FULL: n == NP (even, no padding) -- every per-pair guard folds away at compile time
template <typename T, int NP, bool EVECS, bool FULL>
__global__ void __launch_bounds__(JPD_WARP_THREADS, (WarpLaunch<T, NP>::kMinBlocks))
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

  // ---- block-wide schedule tables, one per pattern
  unsigned* tblA = reinterpret_cast<unsigned*>(smem_raw);
  unsigned* tblB = tblA + Sh::kTblEntries;
  build_schedule<NP>(tblA, m >> 1, 0);
  build_schedule<NP>(tblB, (m >> 1) - 1, 1);
  __syncthreads();

  unsigned char* gbase = smem_raw + Sh::table_bytes() + (size_t)(warp * G + grp) * Sh::group_bytes(sizeof(T));
  T* M = reinterpret_cast<T*>(gbase);
  T* cs = M + Sh::kMElems;
  T* cc = cs + 2 * Sh::kPairsMax;
  T* ss = cc + Sh::kPairsMax;
  T* ev = ss + Sh::kPairsMax;
  int* perm = reinterpret_cast<int*>(ev + NP);

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
    for (int step = 0; step < (FULL ? NP : m); step += 2) {
      warp_step<T, NP, EVECS, 0, FULL>(M, cs, cc, ss, tblA, e, m, sub, n_rot);
      warp_step<T, NP, EVECS, 1, FULL>(M, cs, cc, ss, tblB, e, m, sub, n_rot);
    }
    ++sweeps;
  }
  // with max_num_sweeps == 0 the reference's loop never runs and it reports failure for n > 1,
  // even on a diagonal matrix (:185-186, :214-219)
  const bool converged = (!need && max_num_sweeps > 0) || n == 1;

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
