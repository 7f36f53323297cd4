// jpd_warp4.cuh -- tier 2 for 17 <= n <= 32: one warp per matrix, BLOCK odd-even ordering.
//
// Path being replaced: Jacobi::Diagonalize, /root/reference/include/jacobi_pd.hpp:159-220
// (pivot search :427-468, CalcRot :233-256, ApplyRot :327-393, ApplyRotLeft :409-419,
//  SortRows :477-508).
//
// Why a second warp-per-matrix kernel ("tier 2"): the first one (jpd_warp.cuh, still
// used for n <= 16) sends every entry of M through shared memory once per step of n/2
// rotations and runs at 91 % of the shared-memory pipe with the FP64 pipe half idle.  Here
//  * the positions are grouped in blocks of two and the odd-even (Brent-Luk) exchange ordering
//    runs on BLOCKS: a block step pairs the blocks (J, J+1) = positions q0..q3 and performs the
//    four cross rotations in two steps, (q0,q2)(q1,q3) with exchange and then (q0,q3)(q1,q2)
//    without, which leaves the two blocks exchanged.  One extra step per sweep rotates the
//    pairs inside the blocks.  Every index pair meets exactly once per sweep, the index order is
//    reversed block-wise after a sweep, nothing is addressed by a data-dependent index
//    (tests/test_schedule.py replays the order), and the sweep counts equal the plain odd-even
//    order's (n = 32: 7.2 vs 7.3 sweeps on U(-1,1) matrices);
//  * so a 4x4 tile {block pair P} x {block pair Q} of M is loaded ONCE for two steps, gets its
//    eight rotations (128 FP64 instructions) in registers and is stored once: half the
//    shared-memory traffic per rotation, half the (c,s) fetches, half the barriers;
//  * M is stored tile-major with a stride of 17 elements per tile and the lane that owns a tile
//    is its bank residue (9P+Q) mod 16, so every 64-bit access of a half warp touches 16
//    different banks in both patterns (pattern B's tiles straddle four stored tiles: same
//    residue plus a compile-time offset);
//  * the eigenvector rows are updated with SCALED ("fast") rotations: row p is kept as
//    f_p * e~_p, a rotation costs 2 FMAs per lane and pair instead of 2 multiplications + 2 FMAs
//    (e~_p -= alpha e~_q, e~_q += beta e~_p with alpha = t f_q/f_p, beta = t f_p/f_q; f *= c),
//    the lane that owns the pivot keeps f and 1/f up to date (1/c comes out of the rotation for
//    free) and the scales are folded back into the rows once per sweep, so they stay within
//    [2^-17, 1] and f * (1/f) within a few ulp of 1.
// Convergence rule, return value and sort as jpd_warp.cuh.
#pragma once
#include "jpd_common.cuh"
#include "jpd_warp.cuh"

namespace jpd {

struct Warp4Shape {
  static constexpr int kNP = 32;
  static constexpr int kTileStride = 17;               // elements per stored 4x4 tile (16 + 1: bank skew)
  static constexpr int kIdMul = 9;                      // tile (P,Q) has id 9P+Q: at most two tiles per residue mod 16
  static constexpr int kTileIds = kIdMul * 7 + 7 + 1;   // 71
  static constexpr int kMElems = kTileIds * kTileStride + 1;   // 1208
  static constexpr int kBP = 8;                         // block pairs of pattern A at n = 32
  static constexpr int kCs = 5;                         // (c,s) entries per block pair (4 used: 16-byte bank groups (5P+k) mod 8 differ)
  // per-warp shared memory in elements of T:  M | (c,s)[8][5] | (alpha,beta)[8][5] | (f, 1/f)[32] | evals[32]
  static constexpr int kWarpElems = kMElems + 2 * kCs * kBP + 2 * kCs * kBP + 64 + 32;
  JPD_CX static constexpr size_t warp_bytes(size_t elem) { return ((kWarpElems * elem + 32 * sizeof(int)) + 15) / 16 * 16; }
  JPD_CX static constexpr size_t table_bytes() { return 2 * 32 * sizeof(unsigned); }
  JPD_CX static constexpr size_t block_bytes(size_t elem, int warps) { return table_bytes() + (size_t)warps * warp_bytes(elem); }
  // storage offset of element (i,j), i <= j
  JPD_HD static int addr(int i, int j) { return (kIdMul * (i >> 2) + (j >> 2)) * kTileStride + 4 * (i & 3) + (j & 3); }
};

// offset of element (r,c) of a tile of pattern OFF (0: aligned, 2: shifted by one block) from
// the storage of the aligned tile with the same (P,Q)
template <int OFF>
JPD_CX constexpr int tile_off(int r, int c) {
  return OFF == 0 ? 4 * r + c
                  : ((r >= 2) ? Warp4Shape::kIdMul * Warp4Shape::kTileStride : 0) + ((c >= 2) ? Warp4Shape::kTileStride : 0) +
                        4 * ((r + 2) & 3) + ((c + 2) & 3);
}

// Pivot pair (mpp, mqq, a): skip test (jacobi_pd.hpp:191), CalcRot (:233-256), diagonal through t
// (:337-338), pivot zeroed (:347); EX also exchanges the two positions.  Branch-free so that the
// two independent pivots of a step interleave; returns whether a rotation was needed, otherwise
// (c,s,t,u) = (1,0,0,1).  u = 1/c.
template <typename T, bool EX>
JPD_HD bool pivot4(T& mpp, T& mqq, T& a, T& c, T& s, T& t, T& u) {
  const bool need = pair_needs_rotation(a, mpp, mqq);
  T d = mqq - mpp, as = a;
  const bool d_is_zero = is_zero(d), d_neg = sign_bit(d);
  if (rotation_range_unsafe(d, as)) rescale_for_rotation(d, as);
  rotation_from_scaled(d, as, d_is_zero, d_neg, c, s, t, u);
  c = need ? c : T(1); s = need ? s : T(0); t = need ? t : T(0); u = need ? u : T(1);
  const T g = need ? t * a : T(0);
  const T npp = mpp - g, nqq = mqq + g;
  a = need ? T(0) : a;
  mpp = EX ? nqq : npp;
  mqq = EX ? npp : nqq;
  return need;
}

// B <- G_K B G_L^T with the plain rotation G = [[c, -s], [s, c]] (no exchange)
template <typename T>
JPD_HD void plain_rotate_block(T& b00, T& b01, T& b10, T& b11, T ck, T sk, T cl, T sl) {
  const T u00 = jfma(ck, b00, -(sk * b10)), u01 = jfma(ck, b01, -(sk * b11));
  const T u10 = jfma(sk, b00, ck * b10), u11 = jfma(sk, b01, ck * b11);
  b00 = jfma(cl, u00, -(sl * u01));
  b01 = jfma(sl, u00, cl * u01);
  b10 = jfma(cl, u10, -(sl * u11));
  b11 = jfma(sl, u10, cl * u11);
}
template <typename T, bool EX>
JPD_HD void rotate_block4(T& b00, T& b01, T& b10, T& b11, T ck, T sk, T cl, T sl) {
  if (EX) exchange_rotate_block(b00, b01, b10, b11, ck, sk, cl, sl);
  else plain_rotate_block(b00, b01, b10, b11, ck, sk, cl, sl);
}
template <typename T, bool EX>
JPD_HD void rotate_pair4(T& x, T& y, T c, T s) {
  if (EX) { exchange_rotate_pair(x, y, c, s); }
  else { const T nx = jfma(c, x, -(s * y)); y = jfma(s, x, c * y); x = nx; }
}
// scaled rotation of two eigenvector rows: 2 FMAs (see the file header)
template <typename T, bool EX>
JPD_HD void fast_rotate_rows(T& x, T& y, T alpha, T beta) {
  const T nx = jfma(-alpha, y, x), ny = jfma(beta, x, y);
  x = EX ? ny : nx;
  y = EX ? nx : ny;
}

#if defined(__CUDACC__)

// scale bookkeeping of one pivot (positions p, q of the pair; f, rf in registers of the pivot's lane)
template <typename T, bool EX>
__device__ __forceinline__ void scale_update(T& fp, T& fq, T& rfp, T& rfq, T c, T t, T u, T& alpha, T& beta) {
  alpha = t * (fq * rfp);
  beta = t * (fp * rfq);
  const T nfp = c * fp, nfq = c * fq, nrp = u * rfp, nrq = u * rfq;
  fp = EX ? nfq : nfp; fq = EX ? nfp : nfq;
  rfp = EX ? nrq : nrp; rfq = EX ? nrp : nrq;
}

// MODE 0: block step of pattern A (block pairs at positions 4J..4J+3), 1: pattern B (4J+2..4J+5,
// the two end blocks sit out), 2: the step inside the blocks (pairs (2j, 2j+1), pattern A's tiles).
template <typename T, bool EVECS, int MODE>
__device__ __forceinline__ void warp4_step(T* __restrict__ M, typename Pair2<T>::type* __restrict__ cs,
                                           typename Pair2<T>::type* __restrict__ ab,
                                           typename Pair2<T>::type* __restrict__ fr, const unsigned* __restrict__ tile_tbl,
                                           T (&e)[EVECS ? 32 : 1], int nbp, int lane, int diag_j, int& n_rot) {
  using Sh = Warp4Shape;
  using T2 = typename Pair2<T>::type;
  constexpr int OFF = (MODE == 1) ? 2 : 0;
  constexpr bool TWO = (MODE != 2);            // two rotation steps per tile visit
  const int npairs = (MODE == 1) ? nbp - 1 : nbp;

  // ---- diagonal tiles: the pivots (one lane per block pair)
  if (diag_j >= 0 && diag_j < npairs) {
    T* D = M + (Sh::kIdMul + 1) * diag_j * Sh::kTileStride;
    T d0 = D[tile_off<OFF>(0, 0)], d1 = D[tile_off<OFF>(1, 1)], d2 = D[tile_off<OFF>(2, 2)], d3 = D[tile_off<OFF>(3, 3)];
    T o01 = D[tile_off<OFF>(0, 1)], o02 = D[tile_off<OFF>(0, 2)], o03 = D[tile_off<OFF>(0, 3)];
    T o12 = D[tile_off<OFF>(1, 2)], o13 = D[tile_off<OFF>(1, 3)], o23 = D[tile_off<OFF>(2, 3)];
    const int p0 = 4 * diag_j + OFF;
    T f0 = T(1), f1 = T(1), f2 = T(1), f3 = T(1), r0 = T(1), r1 = T(1), r2 = T(1), r3 = T(1);
    if (EVECS) {   // (f, 1/f) of the four positions
      const T2 s0 = fr[p0], s1 = fr[p0 + 1], s2 = fr[p0 + 2], s3 = fr[p0 + 3];
      f0 = s0.x; r0 = s0.y; f1 = s1.x; r1 = s1.y; f2 = s2.x; r2 = s2.y; f3 = s3.x; r3 = s3.y;
    }
    T ca, sa, ta, ua, cb, sb, tb, ub, al, be;
    T2 v;
    if (TWO) {
      // step 1: (0,2) and (1,3), with exchange
      n_rot += pivot4<T, true>(d0, d2, o02, ca, sa, ta, ua) ? 1 : 0;
      n_rot += pivot4<T, true>(d1, d3, o13, cb, sb, tb, ub) ? 1 : 0;
      rotate_block4<T, true>(o01, o03, o12, o23, ca, sa, cb, sb);       // rows {0,2} x columns {1,3}
      v.x = ca; v.y = sa; cs[Warp4Shape::kCs * diag_j + 0] = v;
      v.x = cb; v.y = sb; cs[Warp4Shape::kCs * diag_j + 1] = v;
      if (EVECS) {
        scale_update<T, true>(f0, f2, r0, r2, ca, ta, ua, al, be); v.x = al; v.y = be; ab[Warp4Shape::kCs * diag_j + 0] = v;
        scale_update<T, true>(f1, f3, r1, r3, cb, tb, ub, al, be); v.x = al; v.y = be; ab[Warp4Shape::kCs * diag_j + 1] = v;
      }
      // step 2: (0,3) and (1,2), no exchange
      n_rot += pivot4<T, false>(d0, d3, o03, ca, sa, ta, ua) ? 1 : 0;
      n_rot += pivot4<T, false>(d1, d2, o12, cb, sb, tb, ub) ? 1 : 0;
      rotate_block4<T, false>(o01, o02, o13, o23, ca, sa, cb, sb);      // rows {0,3} x columns {1,2}
      v.x = ca; v.y = sa; cs[Warp4Shape::kCs * diag_j + 2] = v;
      v.x = cb; v.y = sb; cs[Warp4Shape::kCs * diag_j + 3] = v;
      if (EVECS) {
        scale_update<T, false>(f0, f3, r0, r3, ca, ta, ua, al, be); v.x = al; v.y = be; ab[Warp4Shape::kCs * diag_j + 2] = v;
        scale_update<T, false>(f1, f2, r1, r2, cb, tb, ub, al, be); v.x = al; v.y = be; ab[Warp4Shape::kCs * diag_j + 3] = v;
      }
    } else {
      // inside the blocks: (0,1) and (2,3), no exchange
      n_rot += pivot4<T, false>(d0, d1, o01, ca, sa, ta, ua) ? 1 : 0;
      n_rot += pivot4<T, false>(d2, d3, o23, cb, sb, tb, ub) ? 1 : 0;
      rotate_block4<T, false>(o02, o03, o12, o13, ca, sa, cb, sb);      // rows {0,1} x columns {2,3}
      v.x = ca; v.y = sa; cs[Warp4Shape::kCs * diag_j + 0] = v;
      v.x = cb; v.y = sb; cs[Warp4Shape::kCs * diag_j + 1] = v;
      if (EVECS) {
        scale_update<T, false>(f0, f1, r0, r1, ca, ta, ua, al, be); v.x = al; v.y = be; ab[Warp4Shape::kCs * diag_j + 0] = v;
        scale_update<T, false>(f2, f3, r2, r3, cb, tb, ub, al, be); v.x = al; v.y = be; ab[Warp4Shape::kCs * diag_j + 1] = v;
      }
    }
    D[tile_off<OFF>(0, 0)] = d0; D[tile_off<OFF>(1, 1)] = d1; D[tile_off<OFF>(2, 2)] = d2; D[tile_off<OFF>(3, 3)] = d3;
    D[tile_off<OFF>(0, 1)] = o01; D[tile_off<OFF>(0, 2)] = o02; D[tile_off<OFF>(0, 3)] = o03;
    D[tile_off<OFF>(1, 2)] = o12; D[tile_off<OFF>(1, 3)] = o13; D[tile_off<OFF>(2, 3)] = o23;
    if (EVECS) {
      v.x = f0; v.y = r0; fr[p0] = v;
      v.x = f1; v.y = r1; fr[p0 + 1] = v;
      v.x = f2; v.y = r2; fr[p0 + 2] = v;
      v.x = f3; v.y = r3; fr[p0 + 3] = v;
    }
  }
  __syncwarp();

  // ---- off-diagonal tiles: rows {block pair P} x columns {block pair Q}, P < Q (ApplyRot :350-390)
  {
    const unsigned ent = tile_tbl[lane];
    const int P = (ent >> 16) & 0xffu, Q = ent >> 24;
    if (ent != 0xffffffffu && Q < npairs) {
      T* B = M + (ent & 0xffffu);
      T t00 = B[tile_off<OFF>(0, 0)], t01 = B[tile_off<OFF>(0, 1)], t02 = B[tile_off<OFF>(0, 2)], t03 = B[tile_off<OFF>(0, 3)];
      T t10 = B[tile_off<OFF>(1, 0)], t11 = B[tile_off<OFF>(1, 1)], t12 = B[tile_off<OFF>(1, 2)], t13 = B[tile_off<OFF>(1, 3)];
      T t20 = B[tile_off<OFF>(2, 0)], t21 = B[tile_off<OFF>(2, 1)], t22 = B[tile_off<OFF>(2, 2)], t23 = B[tile_off<OFF>(2, 3)];
      T t30 = B[tile_off<OFF>(3, 0)], t31 = B[tile_off<OFF>(3, 1)], t32 = B[tile_off<OFF>(3, 2)], t33 = B[tile_off<OFF>(3, 3)];
      const T2 p0 = cs[Warp4Shape::kCs * P + 0], p1 = cs[Warp4Shape::kCs * P + 1], q0 = cs[Warp4Shape::kCs * Q + 0], q1 = cs[Warp4Shape::kCs * Q + 1];
      if (TWO) {
        rotate_block4<T, true>(t00, t02, t20, t22, p0.x, p0.y, q0.x, q0.y);
        rotate_block4<T, true>(t01, t03, t21, t23, p0.x, p0.y, q1.x, q1.y);
        rotate_block4<T, true>(t10, t12, t30, t32, p1.x, p1.y, q0.x, q0.y);
        rotate_block4<T, true>(t11, t13, t31, t33, p1.x, p1.y, q1.x, q1.y);
        const T2 p2 = cs[Warp4Shape::kCs * P + 2], p3 = cs[Warp4Shape::kCs * P + 3], q2 = cs[Warp4Shape::kCs * Q + 2], q3 = cs[Warp4Shape::kCs * Q + 3];
        rotate_block4<T, false>(t00, t03, t30, t33, p2.x, p2.y, q2.x, q2.y);
        rotate_block4<T, false>(t01, t02, t31, t32, p2.x, p2.y, q3.x, q3.y);
        rotate_block4<T, false>(t10, t13, t20, t23, p3.x, p3.y, q2.x, q2.y);
        rotate_block4<T, false>(t11, t12, t21, t22, p3.x, p3.y, q3.x, q3.y);
      } else {
        rotate_block4<T, false>(t00, t01, t10, t11, p0.x, p0.y, q0.x, q0.y);
        rotate_block4<T, false>(t02, t03, t12, t13, p0.x, p0.y, q1.x, q1.y);
        rotate_block4<T, false>(t20, t21, t30, t31, p1.x, p1.y, q0.x, q0.y);
        rotate_block4<T, false>(t22, t23, t32, t33, p1.x, p1.y, q1.x, q1.y);
      }
      B[tile_off<OFF>(0, 0)] = t00; B[tile_off<OFF>(0, 1)] = t01; B[tile_off<OFF>(0, 2)] = t02; B[tile_off<OFF>(0, 3)] = t03;
      B[tile_off<OFF>(1, 0)] = t10; B[tile_off<OFF>(1, 1)] = t11; B[tile_off<OFF>(1, 2)] = t12; B[tile_off<OFF>(1, 3)] = t13;
      B[tile_off<OFF>(2, 0)] = t20; B[tile_off<OFF>(2, 1)] = t21; B[tile_off<OFF>(2, 2)] = t22; B[tile_off<OFF>(2, 3)] = t23;
      B[tile_off<OFF>(3, 0)] = t30; B[tile_off<OFF>(3, 1)] = t31; B[tile_off<OFF>(3, 2)] = t32; B[tile_off<OFF>(3, 3)] = t33;
    }
  }
  // ---- pattern B: the two end blocks.  Rows {0,1} meet the column block pair J (lanes 0..),
  // columns {m-2, m-1} meet the row block pair J (lanes 16..): 4x2 values, rotated along the 4.
  if (MODE == 1) {
    const int j = lane & 15;
    if (j < npairs) {
      const bool top = lane < 16;
      const int sb = top ? j * Sh::kTileStride : (Sh::kIdMul * j + nbp - 1) * Sh::kTileStride + 2;
      const int hi = top ? Sh::kTileStride : Sh::kIdMul * Sh::kTileStride;
      const int ka = top ? 1 : 4, ea = top ? 4 : 1;
      T* S = M + sb;
      // k-th position of the block pair: stored at (k+2)&3 of tile j (k < 2) or tile j+1 (k >= 2)
      T* s0 = S + 2 * ka; T* s1 = S + 3 * ka; T* s2 = S + hi; T* s3 = S + hi + ka;
      T a0 = s0[0], a1 = s1[0], a2 = s2[0], a3 = s3[0];
      T b0 = s0[ea], b1 = s1[ea], b2 = s2[ea], b3 = s3[ea];
      const T2 g0 = cs[Warp4Shape::kCs * j + 0], g1 = cs[Warp4Shape::kCs * j + 1], g2 = cs[Warp4Shape::kCs * j + 2], g3 = cs[Warp4Shape::kCs * j + 3];
      rotate_pair4<T, true>(a0, a2, g0.x, g0.y); rotate_pair4<T, true>(b0, b2, g0.x, g0.y);
      rotate_pair4<T, true>(a1, a3, g1.x, g1.y); rotate_pair4<T, true>(b1, b3, g1.x, g1.y);
      rotate_pair4<T, false>(a0, a3, g2.x, g2.y); rotate_pair4<T, false>(b0, b3, g2.x, g2.y);
      rotate_pair4<T, false>(a1, a2, g3.x, g3.y); rotate_pair4<T, false>(b1, b2, g3.x, g3.y);
      s0[0] = a0; s1[0] = a1; s2[0] = a2; s3[0] = a3;
      s0[ea] = b0; s1[ea] = b1; s2[ea] = b2; s3[ea] = b3;
    }
  }
  // ---- eigenvector rows (ApplyRotLeft :414-418), registers only, scaled rotations
  if (EVECS) {
#pragma unroll
    for (int J = 0; J < Sh::kBP - (MODE == 1 ? 1 : 0); ++J) {
      if (J < npairs) {
        constexpr int dummy = 0; (void)dummy;
        const int p = 4 * J + OFF;    // compile-time after unrolling
        const T2 h0 = ab[Warp4Shape::kCs * J + 0], h1 = ab[Warp4Shape::kCs * J + 1];
        if (TWO) {
          fast_rotate_rows<T, true>(e[p], e[p + 2], h0.x, h0.y);
          fast_rotate_rows<T, true>(e[p + 1], e[p + 3], h1.x, h1.y);
          const T2 h2 = ab[Warp4Shape::kCs * J + 2], h3 = ab[Warp4Shape::kCs * J + 3];
          fast_rotate_rows<T, false>(e[p], e[p + 3], h2.x, h2.y);
          fast_rotate_rows<T, false>(e[p + 1], e[p + 2], h3.x, h3.y);
        } else {
          fast_rotate_rows<T, false>(e[p], e[p + 1], h0.x, h0.y);
          fast_rotate_rows<T, false>(e[p + 2], e[p + 3], h1.x, h1.y);
        }
      }
    }
  }
  __syncwarp();
}

#ifndef JPD_WARP4_MINB
#define JPD_WARP4_MINB 3
#endif

template <typename T, bool EVECS>
__global__ void __launch_bounds__(JPD_WARP_THREADS, JPD_WARP4_MINB)
jpd_warp4_kernel(const T* __restrict__ mat, T* __restrict__ evals, T* __restrict__ evecs,
                 int* __restrict__ n_iters, int n, long long batch, int soa, int sort_criteria,
                 int max_num_sweeps) {
  using Sh = Warp4Shape;
  using T2 = typename Pair2<T>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int m = (n + 3) & ~3;          // positions, a multiple of 4 (the padding is zero rows/columns)
  const int nbp = m >> 2;              // block pairs of pattern A
  const int nb = m >> 1;               // blocks = block steps per sweep

  // ---- block-wide tile tables: lane -> tile of pattern A / B (bank residue classes, see header)
  unsigned* tblA = reinterpret_cast<unsigned*>(smem_raw);
  unsigned* tblB = tblA + 32;
  if (threadIdx.x < 64) tblA[threadIdx.x] = 0xffffffffu;
  __syncthreads();
  if (threadIdx.x < 2) {
    unsigned* tbl = threadIdx.x == 0 ? tblA : tblB;
    const int np = threadIdx.x == 0 ? nbp : nbp - 1;
    for (int P = 0; P < np; ++P)
      for (int Q = P + 1; Q < np; ++Q) {
        const int id = Sh::kIdMul * P + Q, cl = id & 15;
        const int slot = (tbl[cl] == 0xffffffffu) ? cl : cl + 16;
        tbl[slot] = (unsigned)(id * Sh::kTileStride) | ((unsigned)P << 16) | ((unsigned)Q << 24);
      }
  }
  __syncthreads();

  unsigned char* wbase = smem_raw + Sh::table_bytes() + (size_t)warp * Sh::warp_bytes(sizeof(T));
  T* M = reinterpret_cast<T*>(wbase);
  T2* cs = reinterpret_cast<T2*>(M + Sh::kMElems);
  T2* ab = cs + Sh::kCs * Sh::kBP;
  T2* fr = ab + Sh::kCs * Sh::kBP;
  T* ev = reinterpret_cast<T*>(fr + 32);
  int* perm = reinterpret_cast<int*>(ev + 32);
  // the lane that owns the diagonal tile J is its bank residue 10J mod 16 (J = 5 (lane/2) mod 8)
  const int diag_j = ((lane & 1) || lane >= 16) ? -1 : ((5 * (lane >> 1)) & 7);

  const long long b = (long long)blockIdx.x * nwarps + warp;
  const bool live = b < batch;

  // ---- load: upper triangle only (jacobi_pd.hpp:167-171); everything else zero
  for (int i = lane; i < Sh::kMElems; i += 32) M[i] = T(0);
  { T2 one; one.x = T(1); one.y = T(1); fr[lane] = one; }
  __syncwarp();
  if (live) {
    for (int r = 0; r < n; ++r) {
      if (lane >= r && lane < n) {
        const T v = soa ? mat[((size_t)r * n + lane) * (size_t)batch + (size_t)b]
                        : mat[(size_t)b * n * n + (size_t)r * n + lane];
        M[Sh::addr(r, lane)] = v;
      }
    }
  }
  T e[EVECS ? 32 : 1];
  if (EVECS) {
#pragma unroll
    for (int k = 0; k < 32; ++k) e[k] = (k == lane) ? T(1) : T(0);   // :172-178
  }
  __syncwarp();

  // ---- sweeps
  int n_rot = 0, sweeps = 0;
  bool need = false;
  for (;;) {
    // the reference's convergence test (:191) on every pair (i, i+d mod m)
    need = false;
    if (lane < m) {
      const T dii = M[Sh::addr(lane, lane)];
      for (int d = 1; d <= (m >> 1); ++d) {
        int j = lane + d; if (j >= m) j -= m;
        const int lo = lane < j ? lane : j, hi = lane < j ? j : lane;
        need = need || pair_needs_rotation(M[Sh::addr(lo, hi)], dii, M[Sh::addr(j, j)]);
      }
    }
    need = __any_sync(0xffffffffu, need);
    if (!need || sweeps >= max_num_sweeps) break;
    warp4_step<T, EVECS, 2>(M, cs, ab, fr, tblA, e, nbp, lane, diag_j, n_rot);
    for (int bs = 0; bs < nb; bs += 2) {
      warp4_step<T, EVECS, 0>(M, cs, ab, fr, tblA, e, nbp, lane, diag_j, n_rot);
      warp4_step<T, EVECS, 1>(M, cs, ab, fr, tblB, e, nbp, lane, diag_j, n_rot);
    }
    ++sweeps;
    // fold the row scales back into the eigenvector rows (keeps f and 1/f near 1)
    if (EVECS) {
#pragma unroll
      for (int k = 0; k < 32; ++k) e[k] *= fr[k].x;
      __syncwarp();
      { T2 one; one.x = T(1); one.y = T(1); fr[lane] = one; }
      __syncwarp();
    }
  }
  const bool converged = (!need && max_num_sweeps > 0) || n == 1;

#pragma unroll
  for (int off = 16; off > 0; off >>= 1) n_rot += __shfl_xor_sync(0xffffffffu, n_rot, off);

  // ---- eigenvalues (:207-209).  A sweep reverses the order of the BLOCKS (the order inside a
  // block is kept), so position p holds original index p after an even number of sweeps and
  // 2 (nb-1 - p/2) + (p&1) after an odd one; the padded indices are those >= n.
  auto orig_index = [&](int p) { return (sweeps & 1) ? 2 * (nb - 1 - (p >> 1)) + (p & 1) : p; };
  const bool real = lane < m && orig_index(lane) < n;
  T my_ev = T(0);
  if (lane < m) { my_ev = M[Sh::addr(lane, lane)]; ev[lane] = my_ev; }
  __syncwarp();
  // sort: output position = number of real slots that precede it (keys as order-preserving
  // integers, ties by slot); same KEY order as SortRows (:477-508)
  if (real) {
    int rank = 0;
    const bool sorted = sort_criteria >= SORT_DECREASING_EVALS && sort_criteria <= SORT_INCREASING_ABS_EVALS;
    const bool use_abs = sort_criteria >= SORT_DECREASING_ABS_EVALS;
    const bool decreasing = (sort_criteria == SORT_DECREASING_EVALS) || (sort_criteria == SORT_DECREASING_ABS_EVALS);
    const long long mine = sortable_bits(use_abs ? jabs(my_ev) : my_ev);
    for (int j = 0; j < m; ++j) {
      if (orig_index(j) >= n) continue;
      bool before;
      if (sorted) {
        const T o = ev[j];
        const long long other = sortable_bits(use_abs ? jabs(o) : o);
        before = (decreasing ? (other > mine) : (other < mine)) || (other == mine && j < lane);
      } else {
        before = j < lane;
      }
      rank += before ? 1 : 0;
    }
    perm[rank] = lane;
  }
  __syncwarp();

  // ---- store (all reads of M are done: its storage now stages the eigenvector rows)
  if (live && lane < n) {
    const T v = ev[perm[lane]];
    if (soa) evals[(size_t)lane * (size_t)batch + (size_t)b] = v;
    else evals[(size_t)b * n + lane] = v;
  }
  if (live && lane == 0 && n_iters) n_iters[b] = converged ? n_rot + 1 : 0;
  if (EVECS) {
    T* Es = M;
#pragma unroll
    for (int k = 0; k < 32; ++k) Es[k * 32 + lane] = e[k];
    __syncwarp();
    if (live && lane < n) {
      for (int k = 0; k < n; ++k) {
        const T v = Es[perm[k] * 32 + lane];
        if (soa) evecs[((size_t)k * n + lane) * (size_t)batch + (size_t)b] = v;
        else evecs[(size_t)b * n * n + (size_t)k * n + lane] = v;
      }
    }
  }
}

#endif  // __CUDACC__

}  // namespace jpd
