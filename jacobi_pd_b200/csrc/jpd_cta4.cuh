// jpd_cta4.cuh -- tier 3 (33 <= n <= 128): one thread block per matrix, BLOCK odd-even ordering,
// 4x4 tiles of M in registers for two rotation steps, scaled rotations for the eigenvectors.
//
// Path being replaced: Jacobi::Diagonalize, /root/reference/include/jacobi_pd.hpp:159-220
// (pivot search :427-468, CalcRot :233-256, ApplyRot :327-393, ApplyRotLeft :409-419,
//  SortRows :477-508).
//
// Why ("tier 3"): the first-generation kernel (jpd_cta.cuh: 2x2 blocks, one rotation
// step per trip of M through shared memory) ran with the shared-memory pipe 73 % busy, the FP64
// pipe 34 % and 1.2 barrier stalls per issue.  Here
//  * the positions are grouped in blocks of two and the odd-even (Brent-Luk) exchange ordering
//    runs on BLOCKS.  A block step pairs the blocks (J, J+1) = positions q0..q3 and does the four
//    cross rotations in two steps: (q0,q2)(q1,q3) with exchange, then (q0,q3)(q1,q2) without,
//    which leaves the two blocks exchanged.  One extra step per sweep rotates the pairs inside
//    the blocks.  Every index pair meets exactly once per sweep; after a sweep the order of the
//    blocks is reversed; nothing is addressed by a data-dependent index (tests/test_schedule.py
//    replays the order); the sweep counts equal the plain odd-even order's.
//  * A 4x4 tile {block pair P} x {block pair Q} of M is therefore loaded ONCE for two steps, gets
//    its eight rotations in registers and is stored once: half the shared-memory traffic, half
//    the (c,s) fetches and half the barriers per rotation.
//  * M is stored tile-major, 17 elements per tile, tile (P,Q) at (P*W + Q)*17: threads that own
//    consecutive tiles touch different banks in both patterns (pattern B's tiles straddle four
//    stored tiles at compile-time offsets).
//  * E (n x n) lives in registers, thread (vp, ech) owns rows 16 ech .. 16 ech + 15 of the TWO columns vp and
//    vp + n/2 (so that every broadcast (alpha, beta) it fetches serves two columns: the fetches were the largest
//    single item of shared-memory traffic with one column of 32 rows per thread), and is updated with SCALED ("fast") rotations: row p is kept as f_p * e~_p, a rotation costs 2
//    FMAs per thread and pair instead of 2 multiplications + 2 FMAs (e~_p -= alpha e~_q,
//    e~_q += beta e~_p, alpha = t f_q/f_p, beta = t f_p/f_q, f *= c).  The thread that owns the
//    pivot keeps f and 1/f (1/c falls out of the rotation), and the scales are folded back into
//    the rows once per sweep, so they stay within [2^-65, 1] and f * (1/f) within a few ulp of 1.
//    In pattern B the block pair that straddles two 16-row chunks is rotated by both owners
//    (each receives the other's two rows through a small shared buffer).
//  * The eigenvector phase of a block step runs beside the pivot phase of the next one (the
//    pivots occupy one warp); (alpha, beta) are double-buffered.
// Convergence rule, return value and sort as jpd_warp.cuh.
#pragma once
#include "jpd_common.cuh"
#include "jpd_warp.cuh"

namespace jpd {

template <int NPB> struct Cta4Shape {
  static constexpr int kChunks = NPB / 32;
  static constexpr int kThreads = NPB * kChunks;
  static constexpr int kBP = NPB / 4;                   // block pairs of pattern A at n = NPB
  static constexpr int kW = ((kBP + 15) / 16) * 16 + 1;  // tile (P,Q) has id P*kW + Q; kW == 1 (mod 16): bank class (P+Q) mod 16
  static constexpr int kTileStride = 17;                // 16 + 1: consecutive tiles start in consecutive banks
  static constexpr int kMElems = ((kBP * kW * kTileStride + 1) / 2) * 2 > NPB * NPB ? ((kBP * kW * kTileStride + 1) / 2) * 2 : NPB * NPB;
  static constexpr int kCs = 5;                         // (c,s) slots per block pair (4 used; 16-byte bank groups (5J+k) mod 8)
  static constexpr int kERows = 16;                     // E: a thread owns rows 16 ech .. 16 ech + 15 of the two columns vp, vp + NPB/2
  static constexpr int kEChunks = NPB / kERows;
  // elements of T:  M | (c,s) | (alpha,beta) x2 | (f,1/f) | evals | boundary rows lo, hi
  static constexpr int kElems = kMElems + 2 * kCs * kBP + 2 * 2 * kCs * kBP + 2 * NPB + NPB + 2 * 2 * kEChunks * NPB;
  static constexpr int kHalfWarps = kThreads / 16;
  static constexpr int kIt = 2;                         // tile slots per thread and pattern (the second one is rarely used)
  JPD_CX static constexpr size_t table_bytes() { return (size_t)2 * kIt * kThreads * sizeof(unsigned); }
  JPD_CX static constexpr size_t block_bytes(size_t elem) {
    return table_bytes() + ((size_t)kElems * elem + 15) / 16 * 16 + (size_t)(NPB + 4) * sizeof(int);
  }
  // storage offset of element (i,j), i <= j
  JPD_HD static int addr(int i, int j) { return ((i >> 2) * kW + (j >> 2)) * kTileStride + 4 * (i & 3) + (j & 3); }
};

// offset of element (r,c) of a tile of pattern OFF (0: aligned, 2: shifted by one block) from the
// storage of the aligned tile with the same (P,Q); W = tiles per stored row
template <int OFF, int W>
JPD_CX constexpr int tile4_off(int r, int c) {
  return OFF == 0 ? 4 * r + c
                  : ((r >= 2) ? W * 17 : 0) + ((c >= 2) ? 17 : 0) + 4 * ((r + 2) & 3) + ((c + 2) & 3);
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
// scale bookkeeping of one pivot: (alpha, beta) for the rows, then f <- c f, 1/f <- (1/c)(1/f)
template <typename T, bool EX>
JPD_HD void scale_update4(T& fp, T& fq, T& rfp, T& rfq, T c, T t, T u, T& alpha, T& beta) {
  alpha = t * (fq * rfp);
  beta = t * (fp * rfq);
  const T nfp = c * fp, nfq = c * fq, nrp = u * rfp, nrq = u * rfq;
  fp = EX ? nfq : nfp; fq = EX ? nfp : nfq;
  rfp = EX ? nrq : nrp; rfq = EX ? nrp : nrq;
}
// the four scaled rotations of a block step on rows x0..x3 (positions q0..q3); h0..h3 = (alpha, beta) of the
// step's rotations (the inner step uses two)
template <typename T, typename T2, int MODE>
JPD_HD void fast_rotate_group(T& x0, T& x1, T& x2, T& x3, T2 h0, T2 h1, T2 h2, T2 h3) {
  if (MODE != 2) {
    fast_rotate_rows<T, true>(x0, x2, h0.x, h0.y);
    fast_rotate_rows<T, true>(x1, x3, h1.x, h1.y);
    fast_rotate_rows<T, false>(x0, x3, h2.x, h2.y);
    fast_rotate_rows<T, false>(x1, x2, h3.x, h3.y);
  } else {
    fast_rotate_rows<T, false>(x0, x1, h0.x, h0.y);
    fast_rotate_rows<T, false>(x2, x3, h1.x, h1.y);
  }
}

#if defined(__CUDACC__)

// shared-memory load that must not be scheduled before `after` has been computed (a fake input operand)
__device__ __forceinline__ double2 lds_after(const double2* p, double after) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"((unsigned)__cvta_generic_to_shared(p)), "d"(after));
  return v;
}
__device__ __forceinline__ double lds_after(const double* p, double after) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(p)), "d"(after));
  return v;
}
__device__ __forceinline__ float lds_after(const float* p, float after) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned)__cvta_generic_to_shared(p)), "f"(after));
  return v;
}
__device__ __forceinline__ float2 lds_after(const float2* p, float after) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"((unsigned)__cvta_generic_to_shared(p)), "f"(after));
  return v;
}

template <typename T> struct Cta4Ctx {
  using T2 = typename Pair2<T>::type;
  T* M; T2* cs; T2* ab; T2* fr; T* xlo; T* xhi;
  const unsigned* tblA; const unsigned* tblB;
  int m, nbp, v, ch;      // v, ch: column / 32-row chunk of the convergence test
  int c0, ech;            // E: first column (the second is c0 + NPB/2) and 16-row chunk
};

// MODE 0: block step of pattern A (block pairs at positions 4J..4J+3), 1: pattern B (4J+2..4J+5,
// the two end blocks sit out), 2: the step inside the blocks (pairs (2j, 2j+1), pattern A's tiles).
// (alpha, beta) buffer of a mode: A -> 1, B and the inner step -> 0 (the steps alternate I A B A B ...).
JPD_CX constexpr int ab_buffer(int mode) { return mode == 0 ? 1 : 0; }

// ---- pivots: thread J < npairs owns the diagonal tile of block pair J
template <typename T, int NPB, bool EVECS, int MODE>
__device__ __forceinline__ void cta4_diag(const Cta4Ctx<T>& cx, int& n_rot) {
  using Sh = Cta4Shape<NPB>;
  using T2 = typename Pair2<T>::type;
  constexpr int OFF = (MODE == 1) ? 2 : 0;
  constexpr int W = Sh::kW;
  const int npairs = (MODE == 1) ? cx.nbp - 1 : cx.nbp;
  const int J = threadIdx.x;
  if (J >= npairs) return;
  T* D = cx.M + (W + 1) * J * Sh::kTileStride;
  T d0 = D[tile4_off<OFF, W>(0, 0)], d1 = D[tile4_off<OFF, W>(1, 1)], d2 = D[tile4_off<OFF, W>(2, 2)], d3 = D[tile4_off<OFF, W>(3, 3)];
  T o01 = D[tile4_off<OFF, W>(0, 1)], o02 = D[tile4_off<OFF, W>(0, 2)], o03 = D[tile4_off<OFF, W>(0, 3)];
  T o12 = D[tile4_off<OFF, W>(1, 2)], o13 = D[tile4_off<OFF, W>(1, 3)], o23 = D[tile4_off<OFF, W>(2, 3)];
  const int p0 = 4 * J + OFF;
  T f0 = T(1), f1 = T(1), f2 = T(1), f3 = T(1), r0 = T(1), r1 = T(1), r2 = T(1), r3 = T(1);
  if (EVECS) {   // (f, 1/f) of the four positions
    const T2 s0 = cx.fr[p0], s1 = cx.fr[p0 + 1], s2 = cx.fr[p0 + 2], s3 = cx.fr[p0 + 3];
    f0 = s0.x; r0 = s0.y; f1 = s1.x; r1 = s1.y; f2 = s2.x; r2 = s2.y; f3 = s3.x; r3 = s3.y;
  }
  T2* cs = cx.cs + Sh::kCs * J;
  T2* ab = cx.ab + ab_buffer(MODE) * Sh::kCs * Sh::kBP + Sh::kCs * J;
  T ca, sa, ta, ua, cb, sb, tb, ub, al, be;
  T2 w;
  if (MODE != 2) {
    // step 1: (0,2) and (1,3), with exchange
    n_rot += pivot4<T, true>(d0, d2, o02, ca, sa, ta, ua) ? 1 : 0;
    n_rot += pivot4<T, true>(d1, d3, o13, cb, sb, tb, ub) ? 1 : 0;
    rotate_block4<T, true>(o01, o03, o12, o23, ca, sa, cb, sb);       // rows {0,2} x columns {1,3}
    w.x = ca; w.y = sa; cs[0] = w;
    w.x = cb; w.y = sb; cs[1] = w;
    if (EVECS) {
      scale_update4<T, true>(f0, f2, r0, r2, ca, ta, ua, al, be); w.x = al; w.y = be; ab[0] = w;
      scale_update4<T, true>(f1, f3, r1, r3, cb, tb, ub, al, be); w.x = al; w.y = be; ab[1] = w;
    }
    // step 2: (0,3) and (1,2), no exchange
    n_rot += pivot4<T, false>(d0, d3, o03, ca, sa, ta, ua) ? 1 : 0;
    n_rot += pivot4<T, false>(d1, d2, o12, cb, sb, tb, ub) ? 1 : 0;
    rotate_block4<T, false>(o01, o02, o13, o23, ca, sa, cb, sb);      // rows {0,3} x columns {1,2}
    w.x = ca; w.y = sa; cs[2] = w;
    w.x = cb; w.y = sb; cs[3] = w;
    if (EVECS) {
      scale_update4<T, false>(f0, f3, r0, r3, ca, ta, ua, al, be); w.x = al; w.y = be; ab[2] = w;
      scale_update4<T, false>(f1, f2, r1, r2, cb, tb, ub, al, be); w.x = al; w.y = be; ab[3] = w;
    }
  } else {
    // inside the blocks: (0,1) and (2,3), no exchange
    n_rot += pivot4<T, false>(d0, d1, o01, ca, sa, ta, ua) ? 1 : 0;
    n_rot += pivot4<T, false>(d2, d3, o23, cb, sb, tb, ub) ? 1 : 0;
    rotate_block4<T, false>(o02, o03, o12, o13, ca, sa, cb, sb);      // rows {0,1} x columns {2,3}
    w.x = ca; w.y = sa; cs[0] = w;
    w.x = cb; w.y = sb; cs[1] = w;
    if (EVECS) {
      scale_update4<T, false>(f0, f1, r0, r1, ca, ta, ua, al, be); w.x = al; w.y = be; ab[0] = w;
      scale_update4<T, false>(f2, f3, r2, r3, cb, tb, ub, al, be); w.x = al; w.y = be; ab[1] = w;
    }
  }
  D[tile4_off<OFF, W>(0, 0)] = d0; D[tile4_off<OFF, W>(1, 1)] = d1; D[tile4_off<OFF, W>(2, 2)] = d2; D[tile4_off<OFF, W>(3, 3)] = d3;
  D[tile4_off<OFF, W>(0, 1)] = o01; D[tile4_off<OFF, W>(0, 2)] = o02; D[tile4_off<OFF, W>(0, 3)] = o03;
  D[tile4_off<OFF, W>(1, 2)] = o12; D[tile4_off<OFF, W>(1, 3)] = o13; D[tile4_off<OFF, W>(2, 3)] = o23;
  if (EVECS) {
    w.x = f0; w.y = r0; cx.fr[p0] = w;
    w.x = f1; w.y = r1; cx.fr[p0 + 1] = w;
    w.x = f2; w.y = r2; cx.fr[p0 + 2] = w;
    w.x = f3; w.y = r3; cx.fr[p0 + 3] = w;
  }
}

// ---- off-diagonal tiles: rows {block pair P} x columns {block pair Q}, P < Q (ApplyRot :350-390),
// and in pattern B the strips of the two end blocks
template <typename T, int NPB, int MODE>
__device__ __forceinline__ void cta4_tiles(const Cta4Ctx<T>& cx) {
  using Sh = Cta4Shape<NPB>;
  using T2 = typename Pair2<T>::type;
  constexpr int OFF = (MODE == 1) ? 2 : 0;
  constexpr int W = Sh::kW;
  const int npairs = (MODE == 1) ? cx.nbp - 1 : cx.nbp;
#pragma unroll 1
  for (int it = 0; it < Sh::kIt; ++it) {
  const unsigned ent = (MODE == 1 ? cx.tblB : cx.tblA)[it * Sh::kThreads + threadIdx.x];
  if (ent == 0xffffffffu) continue;
  {
    T* B = cx.M + (ent & 0xfffffu);
    const T2* csP = cx.cs + Sh::kCs * ((ent >> 20) & 0x3fu);
    const T2* csQ = cx.cs + Sh::kCs * (ent >> 26);
    T t00, t01, t02, t03, t10, t11, t12, t13, t20, t21, t22, t23, t30, t31, t32, t33;
    if (MODE != 2 && sizeof(T) == 8) {
      // fp64 runs at the register cap (64 registers of eigenvector rows): the entries and the (c,s) pairs are loaded
      // group by group, each load pinned behind a result of the group before it (lds_after) -- with everything
      // loaded up front ptxas keeps an eigenvector element in local memory (one LDL + STL per thread and step)
#define JPD_L4(r, c, after) lds_after(B + tile4_off<OFF, W>(r, c), after)
      const T2 p0 = csP[0], q0 = csQ[0];
      t00 = B[tile4_off<OFF, W>(0, 0)]; t02 = B[tile4_off<OFF, W>(0, 2)]; t20 = B[tile4_off<OFF, W>(2, 0)]; t22 = B[tile4_off<OFF, W>(2, 2)];
      rotate_block4<T, true>(t00, t02, t20, t22, p0.x, p0.y, q0.x, q0.y);
      const T2 q1 = lds_after(csQ + 1, t00);
      t01 = JPD_L4(0, 1, t00); t03 = JPD_L4(0, 3, t00); t21 = JPD_L4(2, 1, t00); t23 = JPD_L4(2, 3, t00);
      rotate_block4<T, true>(t01, t03, t21, t23, p0.x, p0.y, q1.x, q1.y);
      const T2 p1 = lds_after(csP + 1, t01);
      t10 = JPD_L4(1, 0, t01); t12 = JPD_L4(1, 2, t01); t30 = JPD_L4(3, 0, t01); t32 = JPD_L4(3, 2, t01);
      rotate_block4<T, true>(t10, t12, t30, t32, p1.x, p1.y, q0.x, q0.y);
      t11 = JPD_L4(1, 1, t10); t13 = JPD_L4(1, 3, t10); t31 = JPD_L4(3, 1, t10); t33 = JPD_L4(3, 3, t10);
      rotate_block4<T, true>(t11, t13, t31, t33, p1.x, p1.y, q1.x, q1.y);
      const T2 p2 = lds_after(csP + 2, t11), q2 = lds_after(csQ + 2, t11);
      rotate_block4<T, false>(t00, t03, t30, t33, p2.x, p2.y, q2.x, q2.y);
      const T2 q3 = lds_after(csQ + 3, t00);
      rotate_block4<T, false>(t01, t02, t31, t32, p2.x, p2.y, q3.x, q3.y);
      const T2 p3 = lds_after(csP + 3, t01);
      rotate_block4<T, false>(t10, t13, t20, t23, p3.x, p3.y, q2.x, q2.y);
      rotate_block4<T, false>(t11, t12, t21, t22, p3.x, p3.y, q3.x, q3.y);
#undef JPD_L4
    } else {
    t00 = B[tile4_off<OFF, W>(0, 0)]; t01 = B[tile4_off<OFF, W>(0, 1)]; t02 = B[tile4_off<OFF, W>(0, 2)]; t03 = B[tile4_off<OFF, W>(0, 3)];
    t10 = B[tile4_off<OFF, W>(1, 0)]; t11 = B[tile4_off<OFF, W>(1, 1)]; t12 = B[tile4_off<OFF, W>(1, 2)]; t13 = B[tile4_off<OFF, W>(1, 3)];
    t20 = B[tile4_off<OFF, W>(2, 0)]; t21 = B[tile4_off<OFF, W>(2, 1)]; t22 = B[tile4_off<OFF, W>(2, 2)]; t23 = B[tile4_off<OFF, W>(2, 3)];
    t30 = B[tile4_off<OFF, W>(3, 0)]; t31 = B[tile4_off<OFF, W>(3, 1)]; t32 = B[tile4_off<OFF, W>(3, 2)]; t33 = B[tile4_off<OFF, W>(3, 3)];
    const T2 p0 = csP[0], p1 = csP[1], q0 = csQ[0], q1 = csQ[1];
    if (MODE != 2) {
      rotate_block4<T, true>(t00, t02, t20, t22, p0.x, p0.y, q0.x, q0.y);
      rotate_block4<T, true>(t01, t03, t21, t23, p0.x, p0.y, q1.x, q1.y);
      rotate_block4<T, true>(t10, t12, t30, t32, p1.x, p1.y, q0.x, q0.y);
      rotate_block4<T, true>(t11, t13, t31, t33, p1.x, p1.y, q1.x, q1.y);
      const T2 p2 = csP[2], p3 = csP[3], q2 = csQ[2], q3 = csQ[3];
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
    }
    B[tile4_off<OFF, W>(0, 0)] = t00; B[tile4_off<OFF, W>(0, 1)] = t01; B[tile4_off<OFF, W>(0, 2)] = t02; B[tile4_off<OFF, W>(0, 3)] = t03;
    B[tile4_off<OFF, W>(1, 0)] = t10; B[tile4_off<OFF, W>(1, 1)] = t11; B[tile4_off<OFF, W>(1, 2)] = t12; B[tile4_off<OFF, W>(1, 3)] = t13;
    B[tile4_off<OFF, W>(2, 0)] = t20; B[tile4_off<OFF, W>(2, 1)] = t21; B[tile4_off<OFF, W>(2, 2)] = t22; B[tile4_off<OFF, W>(2, 3)] = t23;
    B[tile4_off<OFF, W>(3, 0)] = t30; B[tile4_off<OFF, W>(3, 1)] = t31; B[tile4_off<OFF, W>(3, 2)] = t32; B[tile4_off<OFF, W>(3, 3)] = t33;
  }
  }
  // pattern B: rows {0,1} meet the column block pair j (threads 0..31 of the LAST warp), columns
  // {m-2, m-1} meet the row block pair j (threads 32..63 from the end): 4x2 values, rotated along the 4
  if (MODE == 1) {
    const int r = Sh::kThreads - 1 - (int)threadIdx.x;     // the last warps have no tile at n < NPB
    const int j = r & 31;
    if (r < 64 && j < npairs) {
      const bool top = r < 32;
      const int sb = top ? j * Sh::kTileStride : (W * j + cx.nbp - 1) * Sh::kTileStride + 2;
      const int hi = top ? Sh::kTileStride : W * Sh::kTileStride;
      const int ka = top ? 1 : 4, ea = top ? 4 : 1;
      T* S = cx.M + sb;
      // k-th position of the block pair: stored at (k+2)&3 of tile j (k < 2) or of the next tile (k >= 2)
      T* s0 = S + 2 * ka; T* s1 = S + 3 * ka; T* s2 = S + hi; T* s3 = S + hi + ka;
      T a0 = s0[0], a1 = s1[0], a2 = s2[0], a3 = s3[0];
      T b0 = s0[ea], b1 = s1[ea], b2 = s2[ea], b3 = s3[ea];
      const T2* g = cx.cs + Sh::kCs * j;
      const T2 g0 = g[0], g1 = g[1], g2 = g[2], g3 = g[3];
      rotate_pair4<T, true>(a0, a2, g0.x, g0.y); rotate_pair4<T, true>(b0, b2, g0.x, g0.y);
      rotate_pair4<T, true>(a1, a3, g1.x, g1.y); rotate_pair4<T, true>(b1, b3, g1.x, g1.y);
      rotate_pair4<T, false>(a0, a3, g2.x, g2.y); rotate_pair4<T, false>(b0, b3, g2.x, g2.y);
      rotate_pair4<T, false>(a1, a2, g3.x, g3.y); rotate_pair4<T, false>(b1, b2, g3.x, g3.y);
      s0[0] = a0; s1[0] = a1; s2[0] = a2; s3[0] = a3;
      s0[ea] = b0; s1[ea] = b1; s2[ea] = b2; s3[ea] = b3;
    }
  }
}

// ---- eigenvector rows (ApplyRotLeft :414-418) of a block step of mode MODE: registers, scaled
// rotations, and for pattern B the rows received from the neighbouring chunks
template <typename T, int NPB, int MODE>
__device__ __forceinline__ void cta4_evecs(const Cta4Ctx<T>& cx, T (&e)[32]) {
  using Sh = Cta4Shape<NPB>;
  using T2 = typename Pair2<T>::type;
  constexpr int OFF = (MODE == 1) ? 2 : 0;
  constexpr int kGroups = Sh::kERows / 4;     // block pairs per chunk
  const int npairs = (MODE == 1) ? cx.nbp - 1 : cx.nbp;
  const int ech = cx.ech;
  const T2* ab = cx.ab + ab_buffer(MODE) * Sh::kCs * Sh::kBP;
  const int J0 = kGroups * ech;    // block pair of this chunk's first four rows
  if (MODE == 1) {
    // the block pairs that straddle a chunk boundary: both owners rotate all four rows
    if (ech > 0 && J0 - 1 < npairs) {
      const T2* h = ab + Sh::kCs * (J0 - 1);
      const T2 h0 = h[0], h1 = h[1], h2 = h[2], h3 = h[3];
#pragma unroll
      for (int col = 0; col < 2; ++col) {
        const int c = cx.c0 + col * (NPB / 2), cb = Sh::kERows * col;
        T y0 = cx.xhi[(2 * (ech - 1) + 0) * NPB + c], y1 = cx.xhi[(2 * (ech - 1) + 1) * NPB + c];
        T lo0 = e[cb], lo1 = e[cb + 1];
        fast_rotate_group<T, T2, MODE>(y0, y1, lo0, lo1, h0, h1, h2, h3);
        // (the interior groups below do not touch rows 0, 1, 14, 15 in this pattern)
        e[cb] = lo0; e[cb + 1] = lo1;
      }
    }
    if (ech < Sh::kEChunks - 1 && J0 + kGroups - 1 < npairs) {
      const T2* h = ab + Sh::kCs * (J0 + kGroups - 1);
      const T2 h0 = h[0], h1 = h[1], h2 = h[2], h3 = h[3];
#pragma unroll
      for (int col = 0; col < 2; ++col) {
        const int c = cx.c0 + col * (NPB / 2), cb = Sh::kERows * col;
        T y2 = cx.xlo[(2 * (ech + 1) + 0) * NPB + c], y3 = cx.xlo[(2 * (ech + 1) + 1) * NPB + c];
        T hi0 = e[cb + 14], hi1 = e[cb + 15];
        fast_rotate_group<T, T2, MODE>(hi0, hi1, y2, y3, h0, h1, h2, h3);
        e[cb + 14] = hi0; e[cb + 15] = hi1;
      }
    }
  }
#pragma unroll
  for (int jj = 0; jj < kGroups - (MODE == 1 ? 1 : 0); ++jj) {
    if (J0 + jj < npairs) {
      const T2* h = ab + Sh::kCs * (J0 + jj);
      const T2 h0 = h[0], h1 = h[1], h2 = (MODE != 2) ? h[2] : h[0], h3 = (MODE != 2) ? h[3] : h[1];
      const int p = 4 * jj + OFF;   // compile-time after unrolling
#pragma unroll
      for (int col = 0; col < 2; ++col) {
        const int q = Sh::kERows * col + p;
        fast_rotate_group<T, T2, MODE>(e[q], e[q + 1], e[q + 2], e[q + 3], h0, h1, h2, h3);
      }
    }
  }
}

// the rows next to the chunk boundaries, published for the eigenvector phase of a pattern-B step
template <typename T, int NPB>
__device__ __forceinline__ void cta4_publish_rows(const Cta4Ctx<T>& cx, const T (&e)[32]) {
  using Sh = Cta4Shape<NPB>;
#pragma unroll
  for (int col = 0; col < 2; ++col) {
    const int c = cx.c0 + col * (NPB / 2), cb = Sh::kERows * col;
    if (cx.ech > 0) { cx.xlo[(2 * cx.ech + 0) * NPB + c] = e[cb]; cx.xlo[(2 * cx.ech + 1) * NPB + c] = e[cb + 1]; }
    if (cx.ech < Sh::kEChunks - 1) { cx.xhi[(2 * cx.ech + 0) * NPB + c] = e[cb + 14]; cx.xhi[(2 * cx.ech + 1) * NPB + c] = e[cb + 15]; }
  }
}

// first phase of a block step: the pivots of this step (warp 0) beside the eigenvector phase of the previous one
template <typename T, int NPB, bool EVECS, int MODE, int PREV>
__device__ __forceinline__ void cta4_front(const Cta4Ctx<T>& cx, T (&e)[EVECS ? 32 : 1], int& n_rot) {
  using Sh = Cta4Shape<NPB>;
  if (threadIdx.x < 32) cta4_diag<T, NPB, EVECS, MODE>(cx, n_rot);
  if constexpr (EVECS) {
    if constexpr (PREV >= 0) cta4_evecs<T, NPB, PREV>(cx, e);
    if constexpr (MODE == 1) cta4_publish_rows<T, NPB>(cx, e);   // for this step's eigenvector phase
  }
}

template <typename T, int NPB, bool EVECS>
__global__ void __launch_bounds__(Cta4Shape<NPB>::kThreads, (NPB == 64 ? 4 : ((NPB == 96 && sizeof(T) == 4) ? 2 : 1)))
jpd_cta4_kernel(const T* __restrict__ mat, T* __restrict__ evals, T* __restrict__ evecs,
                int* __restrict__ n_iters, int n, long long batch, int soa, int sort_criteria,
                int max_num_sweeps) {
  using Sh = Cta4Shape<NPB>;
  // internal callers (tier 5's capped pivot solves) pass -(cap): an unfinished solve then reports -(rotations + 1)
  const bool raw_count = max_num_sweeps < 0;
  if (raw_count) max_num_sweeps = -max_num_sweeps;
  using T2 = typename Pair2<T>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int v = tid % NPB, ch = tid / NPB;
  const int m = (n + 3) & ~3;          // positions: a multiple of 4 (the padding is zero rows/columns)
  const int nbp = m >> 2, nb = m >> 1;

  // ---- tile tables: the thread that owns a tile has tid % 16 == bank class (P+Q) mod 16 of the tile
  // (tile ids are P*kW + Q with kW == 1 mod 16), tiles of one class go to different half warps: every
  // 64-bit access of a half warp touches 16 different banks.  kIt entries per thread and pattern.
  unsigned* tblA = reinterpret_cast<unsigned*>(smem_raw);
  unsigned* tblB = tblA + Sh::kIt * Sh::kThreads;
  for (int i = tid; i < 2 * Sh::kIt * Sh::kThreads; i += Sh::kThreads) tblA[i] = 0xffffffffu;
  __syncthreads();
  if (tid < 32) {
    const int cl = tid & 15;
    unsigned* tbl = tid < 16 ? tblA : tblB;
    const int np = tid < 16 ? nbp : nbp - 1;
    int slot = 0;
    for (int P = 0; P < np; ++P)
      for (int Q = P + 1 + ((cl - 2 * P - 1) & 15); Q < np; Q += 16) {
        const int idx = (slot / Sh::kHalfWarps) * Sh::kThreads + (slot % Sh::kHalfWarps) * 16 + cl;
        if (idx < Sh::kIt * Sh::kThreads)
          tbl[idx] = (unsigned)((P * Sh::kW + Q) * Sh::kTileStride) | ((unsigned)P << 20) | ((unsigned)Q << 26);
        ++slot;
      }
  }

  T* M = reinterpret_cast<T*>(smem_raw + Sh::table_bytes());
  T2* cs = reinterpret_cast<T2*>(M + Sh::kMElems);
  T2* ab = cs + Sh::kCs * Sh::kBP;
  T2* fr = ab + 2 * Sh::kCs * Sh::kBP;
  T* ev = reinterpret_cast<T*>(fr + NPB);
  T* xlo = ev + NPB;
  T* xhi = xlo + 2 * Sh::kEChunks * NPB;
  int* perm = reinterpret_cast<int*>(smem_raw + Sh::table_bytes() + ((size_t)Sh::kElems * sizeof(T) + 15) / 16 * 16);
  int* rot_total = perm + NPB;
  Cta4Ctx<T> cx;
  cx.M = M; cx.cs = cs; cx.ab = ab; cx.fr = fr; cx.xlo = xlo; cx.xhi = xhi;
  cx.tblA = tblA; cx.tblB = tblB; cx.m = m; cx.nbp = nbp; cx.v = v; cx.ch = ch;
  cx.c0 = tid % (NPB / 2); cx.ech = tid / (NPB / 2);
  __syncthreads();

  for (long long b = blockIdx.x; b < batch; b += gridDim.x) {
    // ---- load: upper triangle only (jacobi_pd.hpp:167-171); everything else zero
    for (int i = tid; i < Sh::kMElems; i += Sh::kThreads) M[i] = T(0);
    if (tid == 0) *rot_total = 0;
    if (tid < NPB) { T2 one; one.x = T(1); one.y = T(1); fr[tid] = one; }
    __syncthreads();
    for (int idx = tid; idx < n * n; idx += Sh::kThreads) {
      const int r = idx / n, c = idx - r * n;
      if (c >= r) {
        const T val = soa ? mat[(size_t)idx * (size_t)batch + (size_t)b] : mat[(size_t)b * n * n + idx];
        M[Sh::addr(r, c)] = val;
      }
    }
    T e[EVECS ? 32 : 1];
    if (EVECS) {
#pragma unroll
      for (int k = 0; k < 32; ++k)   // :172-178
        e[k] = (Sh::kERows * cx.ech + (k & 15) == cx.c0 + (k >> 4) * (NPB / 2)) ? T(1) : T(0);
    }
    __syncthreads();

    // ---- sweeps
    int n_rot = 0, sweeps = 0;
    bool need = false;
    for (;;) {
      // the reference's convergence test (:191) on every pair (i, i+d mod m)
      int mine = 0;
      if (v < m) {
        const T dii = M[Sh::addr(v, v)];
        for (int d = 1 + ch; d <= (m >> 1); d += Sh::kChunks) {
          int j = v + d; if (j >= m) j -= m;
          const int lo = v < j ? v : j, hi = v < j ? j : v;
          mine |= pair_needs_rotation(M[Sh::addr(lo, hi)], dii, M[Sh::addr(j, j)]) ? 1 : 0;
        }
      }
      need = __syncthreads_or(mine) != 0;
      if (!need || sweeps >= max_num_sweeps) break;
      // the step inside the blocks, then nb block steps A B A B ...; the eigenvector phase of a
      // step runs in the first phase of the next one
      cta4_front<T, NPB, EVECS, 2, -1>(cx, e, n_rot);
      __syncthreads();
      cta4_tiles<T, NPB, 2>(cx);
      __syncthreads();
      for (int bs = 0; bs < nb; bs += 2) {
        if (bs == 0) cta4_front<T, NPB, EVECS, 0, 2>(cx, e, n_rot);
        else cta4_front<T, NPB, EVECS, 0, 1>(cx, e, n_rot);
        __syncthreads();
        cta4_tiles<T, NPB, 0>(cx);
        __syncthreads();
        cta4_front<T, NPB, EVECS, 1, 0>(cx, e, n_rot);
        __syncthreads();
        cta4_tiles<T, NPB, 1>(cx);
        __syncthreads();
      }
      ++sweeps;
      if constexpr (EVECS) {
        cta4_evecs<T, NPB, 1>(cx, e);     // the last B step's eigenvector phase
        // fold the row scales back into the rows (keeps f and 1/f near 1)
#pragma unroll
        for (int k = 0; k < Sh::kERows; ++k) { const T fk = fr[Sh::kERows * cx.ech + k].x; e[k] *= fk; e[Sh::kERows + k] *= fk; }
        __syncthreads();
        if (tid < NPB) { T2 one; one.x = T(1); one.y = T(1); fr[tid] = one; }
      }
    }
    // max_num_sweeps == 0: the reference reports failure for n > 1 whatever the matrix (:185-186, :214-219)
    const bool converged = !need && max_num_sweeps > 0;
    if (n_rot) atomicAdd(rot_total, n_rot);

    // ---- eigenvalues (:207-209).  A sweep reverses the order of the BLOCKS (the order inside a
    // block is kept): position p holds original index p after an even number of sweeps and
    // 2 (nb-1 - p/2) + (p&1) after an odd one; the padded indices are those >= n.
    auto orig_index = [&](int p) { return (sweeps & 1) ? 2 * (nb - 1 - (p >> 1)) + (p & 1) : p; };
    const bool real = tid < m && orig_index(tid) < n;
    T my_ev = T(0);
    if (tid < m) { my_ev = M[Sh::addr(tid, tid)]; ev[tid] = my_ev; }
    __syncthreads();
    if (real) {
      int rank = 0;
      const bool sorted = sort_criteria >= SORT_DECREASING_EVALS && sort_criteria <= SORT_INCREASING_ABS_EVALS;
      const bool use_abs = sort_criteria >= SORT_DECREASING_ABS_EVALS;
      const bool decreasing = (sort_criteria == SORT_DECREASING_EVALS) || (sort_criteria == SORT_DECREASING_ABS_EVALS);
      const long long key = sortable_bits(use_abs ? jabs(my_ev) : my_ev);
      for (int j = 0; j < m; ++j) {
        if (orig_index(j) >= n) continue;
        bool before = j < tid;
        if (sorted) {
          const T o = ev[j];
          const long long other = sortable_bits(use_abs ? jabs(o) : o);
          before = (decreasing ? (other > key) : (other < key)) || (other == key && j < tid);
        }
        rank += before ? 1 : 0;
      }
      perm[rank] = tid;
    }
    __syncthreads();

    // ---- store (M's storage now stages the eigenvector rows)
    if (tid < n) {
      const T val = ev[perm[tid]];
      if (soa) evals[(size_t)tid * (size_t)batch + (size_t)b] = val;
      else evals[(size_t)b * n + tid] = val;
    }
    if (tid == 0 && n_iters) n_iters[b] = converged ? *rot_total + 1 : (raw_count ? -(*rot_total + 1) : 0);
    if (EVECS) {
      T* Es = M;
#pragma unroll
      for (int k = 0; k < 32; ++k) Es[(Sh::kERows * cx.ech + (k & 15)) * NPB + cx.c0 + (k >> 4) * (NPB / 2)] = e[k];
      __syncthreads();
      for (int idx = tid; idx < n * n; idx += Sh::kThreads) {
        const int k = idx / n, c = idx - k * n;
        const T val = Es[perm[k] * NPB + c];
        if (soa) evecs[(size_t)idx * (size_t)batch + (size_t)b] = val;
        else evecs[(size_t)b * n * n + idx] = val;
      }
    }
    __syncthreads();   // M / ev / perm are reused by the next matrix of this block
  }
}

#endif  // __CUDACC__

}  // namespace jpd
