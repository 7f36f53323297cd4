// jpd_cta5.cuh -- tier 3, block odd-even ordering (jpd_cta4.cuh) with the pivot phase FUSED into the
// tile phase: one barrier per block step (two rotation steps).
//
// Path being replaced: Jacobi::Diagonalize, /root/reference/include/jacobi_pd.hpp:159-220
// (CalcRot :233-256, ApplyRot :327-393, ApplyRotLeft :409-419, SortRows :477-508).
//
// In jpd_cta4.cuh a block step is  pivots | barrier | tiles | barrier,  and the profile
// (profiles/r2_tier3_ncu_summary.md) shows more than half of the warp time in the pivot phase: one warp
// walks two dependent rotation chains (~1000 cycles) and then its own eigenvector rows while fifteen
// warps wait.  The pivots of the NEXT block step, however, depend on the tile phase only through the
// super-diagonal tiles: the diagonal 4x4 tile of block pair P in the next pattern consists of pieces of
// the two diagonal tiles next to it (final since the previous pivots) and the bottom-left 2x2 quadrant
// of the current pattern's tile (P, P+1) -- or, at the two ends of pattern B, of the strip next to the
// diagonal.  So the thread that owns that tile keeps the quadrant in registers and goes straight on to the
// next pivots; everybody else goes on to the eigenvector rows of the CURRENT step ((alpha, beta) were
// published one phase earlier).  One barrier per block step; (c,s) and (alpha,beta) double-buffered.
#pragma once
#include "jpd_cta4.cuh"

namespace jpd {

template <int NPB> struct Cta5Shape {
  using B4 = Cta4Shape<NPB>;
  static constexpr int kThreads = B4::kThreads;
  static constexpr int kW = B4::kW;
  static constexpr int kCs = B4::kCs;
  static constexpr int kBP = B4::kBP;
  static constexpr int kChunks = B4::kChunks;
  static constexpr int kTileStride = B4::kTileStride;
  static constexpr int kMElems = B4::kMElems;
  static constexpr int kOvf = 64;      // overflow list of the table builder
  // elements of T:  M | (c,s) x2 | (alpha,beta) x2 | (f,1/f) | evals | boundary rows lo, hi
  static constexpr int kElems = kMElems + 2 * 2 * kCs * kBP + 2 * 2 * kCs * kBP + 2 * NPB + NPB + 2 * 2 * kChunks * NPB;
  static constexpr int kInts = NPB + 4 + 2 * (kOvf + 1);
  JPD_CX static constexpr size_t table_bytes() { return (size_t)2 * kThreads * sizeof(unsigned); }
  JPD_CX static constexpr size_t block_bytes(size_t elem) {
    return table_bytes() + ((size_t)kElems * elem + 15) / 16 * 16 + (size_t)kInts * sizeof(int);
  }
  JPD_HD static int addr(int i, int j) { return B4::addr(i, j); }
  // first half warp that may own table tiles: the lanes 0 .. nbp-1 of warp 0 own the tiles next to the diagonal
  JPD_HD static int first_half_warp(int nbp) { return (nbp + 15) / 16; }
};

#if defined(__CUDACC__)

template <typename T> struct Cta5Ctx {
  using T2 = typename Pair2<T>::type;
  T* M; T2* cs; T2* ab; T2* fr; T* xlo; T* xhi;      // cs, ab: two buffers each (ab_buffer(mode))
  const unsigned* tblA; const unsigned* tblB;
  int m, nbp, v, ch;
};

// ---- pivots of block pair J in pattern MODE.  REG_CROSS: the four entries rows {0,1} x columns {2,3} of the
// diagonal tile come from the caller's registers (x02, x03, x12, x13), everything else from shared memory.
template <typename T, int NPB, bool EVECS, int MODE, bool REG_CROSS>
__device__ __forceinline__ void cta5_diag(const Cta5Ctx<T>& cx, int J, T x02, T x03, T x12, T x13, int& n_rot) {
  using Sh = Cta5Shape<NPB>;
  using T2 = typename Pair2<T>::type;
  constexpr int OFF = (MODE == 1) ? 2 : 0;
  constexpr int W = Sh::kW;
  T* D = cx.M + (W + 1) * J * Sh::kTileStride;
  T d0 = D[tile4_off<OFF, W>(0, 0)], d1 = D[tile4_off<OFF, W>(1, 1)], d2 = D[tile4_off<OFF, W>(2, 2)], d3 = D[tile4_off<OFF, W>(3, 3)];
  T o01 = D[tile4_off<OFF, W>(0, 1)], o23 = D[tile4_off<OFF, W>(2, 3)];
  T o02 = x02, o03 = x03, o12 = x12, o13 = x13;
  if (!REG_CROSS) {
    o02 = D[tile4_off<OFF, W>(0, 2)]; o03 = D[tile4_off<OFF, W>(0, 3)];
    o12 = D[tile4_off<OFF, W>(1, 2)]; o13 = D[tile4_off<OFF, W>(1, 3)];
  }
  const int p0 = 4 * J + OFF;
  T f0 = T(1), f1 = T(1), f2 = T(1), f3 = T(1), r0 = T(1), r1 = T(1), r2 = T(1), r3 = T(1);
  if (EVECS) {   // (f, 1/f) of the four positions
    const T2 s0 = cx.fr[p0], s1 = cx.fr[p0 + 1], s2 = cx.fr[p0 + 2], s3 = cx.fr[p0 + 3];
    f0 = s0.x; r0 = s0.y; f1 = s1.x; r1 = s1.y; f2 = s2.x; r2 = s2.y; f3 = s3.x; r3 = s3.y;
  }
  T2* cs = cx.cs + ab_buffer(MODE) * Sh::kCs * Sh::kBP + Sh::kCs * J;
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

// the eight rotations (two steps) of pattern MODE on a 4x4 tile t[4r+c]; csP, csQ: the (c,s) of the row / column block pair
template <typename T, int MODE>
__device__ __forceinline__ void tile5_rotate(T (&t)[16], const typename Pair2<T>::type* csP, const typename Pair2<T>::type* csQ) {
  using T2 = typename Pair2<T>::type;
  const T2 p0 = csP[0], p1 = csP[1], q0 = csQ[0], q1 = csQ[1];
  if (MODE != 2) {
    rotate_block4<T, true>(t[0], t[2], t[8], t[10], p0.x, p0.y, q0.x, q0.y);
    rotate_block4<T, true>(t[1], t[3], t[9], t[11], p0.x, p0.y, q1.x, q1.y);
    rotate_block4<T, true>(t[4], t[6], t[12], t[14], p1.x, p1.y, q0.x, q0.y);
    rotate_block4<T, true>(t[5], t[7], t[13], t[15], p1.x, p1.y, q1.x, q1.y);
    const T2 p2 = csP[2], p3 = csP[3], q2 = csQ[2], q3 = csQ[3];
    rotate_block4<T, false>(t[0], t[3], t[12], t[15], p2.x, p2.y, q2.x, q2.y);
    rotate_block4<T, false>(t[1], t[2], t[13], t[14], p2.x, p2.y, q3.x, q3.y);
    rotate_block4<T, false>(t[4], t[7], t[8], t[11], p3.x, p3.y, q2.x, q2.y);
    rotate_block4<T, false>(t[5], t[6], t[9], t[10], p3.x, p3.y, q3.x, q3.y);
  } else {
    rotate_block4<T, false>(t[0], t[1], t[4], t[5], p0.x, p0.y, q0.x, q0.y);
    rotate_block4<T, false>(t[2], t[3], t[6], t[7], p0.x, p0.y, q1.x, q1.y);
    rotate_block4<T, false>(t[8], t[9], t[12], t[13], p1.x, p1.y, q0.x, q0.y);
    rotate_block4<T, false>(t[10], t[11], t[14], t[15], p1.x, p1.y, q1.x, q1.y);
  }
}
template <typename T, int OFF, int W>
__device__ __forceinline__ void tile5_load(const T* B, T (&t)[16]) {
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) t[4 * r + c] = B[tile4_off<OFF, W>(r, c)];
}
// SKIP_CROSS: leave out rows {2,3} x columns {0,1} (they go on into the next pivots and are stored there)
template <typename T, int OFF, int W, bool SKIP_CROSS>
__device__ __forceinline__ void tile5_store(T* B, const T (&t)[16]) {
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (!(SKIP_CROSS && r >= 2 && c < 2)) B[tile4_off<OFF, W>(r, c)] = t[4 * r + c];
}

// a strip of pattern B: positions k = 0..3 of block pair j against the two rows 0,1 (top) or the two
// columns m-2, m-1 (right); a[k], b[k] = the two lines.  Pointers to the 8 entries in p[2k+e].
template <typename T, int NPB>
__device__ __forceinline__ void strip5_pointers(const Cta5Ctx<T>& cx, int j, bool top, T* (&p)[8]) {
  using Sh = Cta5Shape<NPB>;
  const int sb = top ? j * Sh::kTileStride : (Sh::kW * j + cx.nbp - 1) * Sh::kTileStride + 2;
  const int hi = top ? Sh::kTileStride : Sh::kW * Sh::kTileStride;
  const int ka = top ? 1 : 4, ea = top ? 4 : 1;
  T* S = cx.M + sb;
  T* s[4] = {S + 2 * ka, S + 3 * ka, S + hi, S + hi + ka};
#pragma unroll
  for (int k = 0; k < 4; ++k) { p[2 * k] = s[k]; p[2 * k + 1] = s[k] + ea; }
}
template <typename T>
__device__ __forceinline__ void strip5_rotate(T (&a)[4], T (&b)[4], const typename Pair2<T>::type* g) {
  using T2 = typename Pair2<T>::type;
  const T2 g0 = g[0], g1 = g[1], g2 = g[2], g3 = g[3];
  rotate_pair4<T, true>(a[0], a[2], g0.x, g0.y); rotate_pair4<T, true>(b[0], b[2], g0.x, g0.y);
  rotate_pair4<T, true>(a[1], a[3], g1.x, g1.y); rotate_pair4<T, true>(b[1], b[3], g1.x, g1.y);
  rotate_pair4<T, false>(a[0], a[3], g2.x, g2.y); rotate_pair4<T, false>(b[0], b[3], g2.x, g2.y);
  rotate_pair4<T, false>(a[1], a[2], g3.x, g3.y); rotate_pair4<T, false>(b[1], b[2], g3.x, g3.y);
}

// One block step of pattern CUR (0: A, 1: B, 2: the step inside the blocks), followed by a step of pattern
// NEXT (-1: none, the sweep ends):  tiles and strips of CUR; the threads next to the diagonal go straight
// on to the pivots of NEXT; eigenvector rows of CUR; one barrier.
template <typename T, int NPB, bool EVECS, bool XW, int CUR, int NEXT>
__device__ __forceinline__ void cta5_phase(const Cta5Ctx<T>& cx, T (&e)[EVECS ? 32 : 1], int& n_rot) {
  using Sh = Cta5Shape<NPB>;
  using T2 = typename Pair2<T>::type;
  constexpr int OFF = (CUR == 1) ? 2 : 0;
  constexpr int W = Sh::kW;
  const int nbp = cx.nbp;
  const int npairs = (CUR == 1) ? nbp - 1 : nbp;
  const T2* cs = cx.cs + ab_buffer(CUR) * Sh::kCs * Sh::kBP;
  // XW: the threads next to the diagonal are an extra warp of their own (no eigenvector rows, no table tile)
  const int tid = XW ? (int)threadIdx.x - Sh::kThreads : (int)threadIdx.x;   // role index
  const int wid = (int)threadIdx.x;                                           // worker index (valid when < kThreads)

  // ---- eigenvector rows of THIS step (ApplyRotLeft :414-418), registers only, fed by broadcast loads; before a
  // pattern-B step the rows next to the chunk boundaries are published for it.  Every warp except the one
  // that holds the threads next to the diagonal does this FIRST: those threads then find an empty
  // shared-memory queue for their tile, and their pivots -- the longest dependent chain of the step --
  // start early and run beside the other warps' tiles.
  auto rows = [&]() {
    if constexpr (EVECS) {
      if (wid < Sh::kThreads) {
        cta4_evecs<T, NPB, CUR>(cx, e);
        if constexpr (NEXT == 1) {
          if (cx.ch > 0) { cx.xlo[(2 * cx.ch + 0) * NPB + cx.v] = e[0]; cx.xlo[(2 * cx.ch + 1) * NPB + cx.v] = e[1]; }
          if (cx.ch < Sh::kChunks - 1) { cx.xhi[(2 * cx.ch + 0) * NPB + cx.v] = e[30]; cx.xhi[(2 * cx.ch + 1) * NPB + cx.v] = e[31]; }
        }
      }
    }
  };
  const bool late_rows = !XW && wid < 32;     // warp 0 holds the threads next to the diagonal
  if (!late_rows) rows();

  if (tid >= 0 && tid < nbp) {
    // ---- the tiles / strips next to the diagonal (lanes of warp 0), then the pivots of NEXT
    if (CUR != 1) {
      if (tid <= npairs - 2) {                       // tile (P, P+1)
        const int P = tid;
        T* B = cx.M + (P * W + P + 1) * Sh::kTileStride;
        T t[16];
        tile5_load<T, OFF, W>(B, t);
        tile5_rotate<T, CUR>(t, cs + Sh::kCs * P, cs + Sh::kCs * (P + 1));
        if (NEXT == 1) {
          tile5_store<T, OFF, W, true>(B, t);
          cta5_diag<T, NPB, EVECS, 1, true>(cx, P, t[8], t[9], t[12], t[13], n_rot);
        } else {
          tile5_store<T, OFF, W, false>(B, t);
        }
      }
      if (CUR == 2 && NEXT == 0)                     // after the inner step the aligned pivots need nothing from the tiles
        cta5_diag<T, NPB, EVECS, 0, false>(cx, tid, T(0), T(0), T(0), T(0), n_rot);
    } else {
      if (tid >= 1 && tid <= npairs - 1) {           // tile (P, P+1) of pattern B, P = tid - 1: feeds the aligned pivots of block pair tid
        const int P = tid - 1;
        T* B = cx.M + (P * W + P + 1) * Sh::kTileStride;
        T t[16];
        tile5_load<T, OFF, W>(B, t);
        tile5_rotate<T, CUR>(t, cs + Sh::kCs * P, cs + Sh::kCs * (P + 1));
        if (NEXT == 0) {
          tile5_store<T, OFF, W, true>(B, t);
          cta5_diag<T, NPB, EVECS, 0, true>(cx, tid, t[8], t[9], t[12], t[13], n_rot);
        } else {
          tile5_store<T, OFF, W, false>(B, t);
        }
      } else if (tid == 0 || tid == npairs) {        // the strips next to the diagonal: top strip 0, right strip npairs-1
        const bool top = tid == 0;
        const int j = top ? 0 : npairs - 1;
        T* p[8];
        strip5_pointers<T, NPB>(cx, j, top, p);
        T a[4], b[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { a[k] = *p[2 * k]; b[k] = *p[2 * k + 1]; }
        strip5_rotate<T>(a, b, cs + Sh::kCs * j);
        if (NEXT == 0) {
          // top: rows 0,1 x columns 2,3 of the first diagonal tile are (a0,a1 / b0,b1); right: rows m-4,m-3 x columns m-2,m-1 are (a2,b2 / a3,b3)
          if (top) { *p[4] = a[2]; *p[5] = b[2]; *p[6] = a[3]; *p[7] = b[3]; }
          else     { *p[0] = a[0]; *p[1] = b[0]; *p[2] = a[1]; *p[3] = b[1]; }
          const T x02 = top ? a[0] : a[2], x03 = top ? a[1] : b[2], x12 = top ? b[0] : a[3], x13 = top ? b[1] : b[3];
          cta5_diag<T, NPB, EVECS, 0, true>(cx, top ? 0 : nbp - 1, x02, x03, x12, x13, n_rot);
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) { *p[2 * k] = a[k]; *p[2 * k + 1] = b[k]; }
        }
      }
    }
  } else if (wid < Sh::kThreads) {
    // ---- the other tiles: rows {block pair P} x columns {block pair Q}, Q > P + 1 (ApplyRot :350-390)
    const unsigned ent = (CUR == 1 ? cx.tblB : cx.tblA)[wid];
    if (ent != 0xffffffffu) {
      T* B = cx.M + (ent & 0xfffffu);
      T t[16];
      tile5_load<T, OFF, W>(B, t);
      tile5_rotate<T, CUR>(t, cs + Sh::kCs * ((ent >> 20) & 0x3fu), cs + Sh::kCs * (ent >> 26));
      tile5_store<T, OFF, W, false>(B, t);
    }
    // pattern B: the other strips (threads counted from the end)
    if (CUR == 1) {
      const int r = Sh::kThreads - 1 - wid;
      const int j = r & 31;
      const bool top = r < 32;
      if (r < 64 && j < npairs && !(top && j == 0) && !(!top && j == npairs - 1)) {
        T* p[8];
        strip5_pointers<T, NPB>(cx, j, top, p);
        T a[4], b[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { a[k] = *p[2 * k]; b[k] = *p[2 * k + 1]; }
        strip5_rotate<T>(a, b, cs + Sh::kCs * j);
#pragma unroll
        for (int k = 0; k < 4; ++k) { *p[2 * k] = a[k]; *p[2 * k + 1] = b[k]; }
      }
    }
  }
  // ---- the warp of the threads next to the diagonal does its eigenvector rows last
  if (late_rows) rows();
  __syncthreads();
}

#define JPD_CTA5_KERNEL_ARGS const T* __restrict__ mat, T* __restrict__ evals, T* __restrict__ evecs, \
                int* __restrict__ n_iters, int n, long long batch, int soa, int sort_criteria, int max_num_sweeps
template <typename T, int NPB, bool EVECS, bool XW>
__device__ __forceinline__ void cta5_body(JPD_CTA5_KERNEL_ARGS);

template <typename T, int NPB, bool EVECS>
__global__ void __launch_bounds__(Cta5Shape<NPB>::kThreads, (NPB == 64 ? 4 : ((NPB == 96 && sizeof(T) == 4) ? 2 : 1)))
jpd_cta5_kernel(JPD_CTA5_KERNEL_ARGS) {
  cta5_body<T, NPB, EVECS, false>(mat, evals, evecs, n_iters, n, batch, soa, sort_criteria, max_num_sweeps);
}
// 17 warps: one scheduler hosts five of them, so a thread gets at most 16384 / (5 * 32) = 102 -> 96 registers
template <typename T, int NPB, bool EVECS>
__global__ void __launch_bounds__(Cta5Shape<NPB>::kThreads + 32, 1)
jpd_cta5x_kernel(JPD_CTA5_KERNEL_ARGS) {
  cta5_body<T, NPB, EVECS, true>(mat, evals, evecs, n_iters, n, batch, soa, sort_criteria, max_num_sweeps);
}

template <typename T, int NPB, bool EVECS, bool XW>
__device__ __forceinline__ void cta5_body(JPD_CTA5_KERNEL_ARGS) {
  using Sh = Cta5Shape<NPB>;
  using T2 = typename Pair2<T>::type;
  // internal callers (tier 5's capped pivot solves) pass -(cap): an unfinished solve then reports -(rotations + 1)
  const bool raw_count = max_num_sweeps < 0;
  if (raw_count) max_num_sweeps = -max_num_sweeps;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const bool worker = tid < Sh::kThreads;          // XW: the last warp only does the pivots
  const int rid = XW ? tid - Sh::kThreads : tid;   // role index of cta5_phase
  const int v = tid % NPB, ch = tid / NPB;
  const int m = (n + 3) & ~3;          // positions: a multiple of 4 (the padding is zero rows/columns)
  const int nbp = m >> 2, nb = m >> 1;

  unsigned* tblA = reinterpret_cast<unsigned*>(smem_raw);
  unsigned* tblB = tblA + Sh::kThreads;
  T* M = reinterpret_cast<T*>(smem_raw + Sh::table_bytes());
  T2* cs = reinterpret_cast<T2*>(M + Sh::kMElems);
  T2* ab = cs + 2 * Sh::kCs * Sh::kBP;
  T2* fr = ab + 2 * Sh::kCs * Sh::kBP;
  T* ev = reinterpret_cast<T*>(fr + NPB);
  T* xlo = ev + NPB;
  T* xhi = xlo + 2 * Sh::kChunks * NPB;
  int* perm = reinterpret_cast<int*>(smem_raw + Sh::table_bytes() + ((size_t)Sh::kElems * sizeof(T) + 15) / 16 * 16);
  int* rot_total = perm + NPB;
  int* ovf = perm + NPB + 4;           // [pattern][count, kOvf entries]

  // ---- tile tables, one entry per thread and pattern.  The tiles (P, P+1) and the lanes 0 .. nbp-1 of warp 0
  // belong together (cta5_phase); every other tile goes to a thread with tid % 16 == its bank class
  // (P+Q) mod 16 in a half warp of its own (conflict-free 64-bit accesses); the few tiles of an over-full
  // class take whatever slot is left (a two-way conflict in that half warp).
  for (int i = tid; i < 2 * Sh::kThreads; i += blockDim.x) tblA[i] = 0xffffffffu;
  if (tid < 2) ovf[tid * (Sh::kOvf + 1)] = 0;
  __syncthreads();
  const int hw0 = XW ? 0 : Sh::first_half_warp(nbp), hws = Sh::kThreads / 16;
  if (tid < 32) {
    const int cl = tid & 15, pat = tid >> 4;
    unsigned* tbl = pat == 0 ? tblA : tblB;
    const int np = pat == 0 ? nbp : nbp - 1;
    int slot = hw0;
    for (int P = 0; P < np; ++P)
      for (int Q = P + 1 + ((cl - 2 * P - 1) & 15); Q < np; Q += 16) {
        if (Q == P + 1) continue;
        const unsigned ent = (unsigned)((P * Sh::kW + Q) * Sh::kTileStride) | ((unsigned)P << 20) | ((unsigned)Q << 26);
        if (slot < hws) { tbl[slot * 16 + cl] = ent; ++slot; }
        else { const int k = atomicAdd(&ovf[pat * (Sh::kOvf + 1)], 1); if (k < Sh::kOvf) ovf[pat * (Sh::kOvf + 1) + 1 + k] = (int)ent; }
      }
  }
  __syncthreads();
  if (tid < 2) {
    unsigned* tbl = tid == 0 ? tblA : tblB;
    const int cnt = min(ovf[tid * (Sh::kOvf + 1)], Sh::kOvf);
    int free_slot = hw0 * 16;
    for (int k = 0; k < cnt; ++k) {
      while (free_slot < Sh::kThreads && tbl[free_slot] != 0xffffffffu) ++free_slot;
      if (free_slot < Sh::kThreads) tbl[free_slot++] = (unsigned)ovf[tid * (Sh::kOvf + 1) + 1 + k];
    }
  }

  Cta5Ctx<T> cx;
  cx.M = M; cx.cs = cs; cx.ab = ab; cx.fr = fr; cx.xlo = xlo; cx.xhi = xhi;
  cx.tblA = tblA; cx.tblB = tblB; cx.m = m; cx.nbp = nbp; cx.v = v; cx.ch = ch;
  __syncthreads();

  for (long long b = blockIdx.x; b < batch; b += gridDim.x) {
    // ---- load: upper triangle only (jacobi_pd.hpp:167-171); everything else zero
    for (int i = tid; i < Sh::kMElems; i += blockDim.x) M[i] = T(0);
    if (tid == 0) *rot_total = 0;
    if (tid < NPB) { T2 one; one.x = T(1); one.y = T(1); fr[tid] = one; }
    __syncthreads();
    for (int idx = tid; idx < n * n; idx += blockDim.x) {
      const int r = idx / n, c = idx - r * n;
      if (c >= r) {
        const T val = soa ? mat[(size_t)idx * (size_t)batch + (size_t)b] : mat[(size_t)b * n * n + idx];
        M[Sh::addr(r, c)] = val;
      }
    }
    T e[EVECS ? 32 : 1];
    if (EVECS) {
#pragma unroll
      for (int k = 0; k < 32; ++k) e[k] = (worker && 32 * ch + k == v) ? T(1) : T(0);   // :172-178
    }
    __syncthreads();

    // ---- sweeps
    int n_rot = 0, sweeps = 0;
    bool need = false;
    for (;;) {
      // the reference's convergence test (:191) on every pair (i, i+d mod m)
      int mine = 0;
      if (worker && v < m) {
        const T dii = M[Sh::addr(v, v)];
        for (int d = 1 + ch; d <= (m >> 1); d += Sh::kChunks) {
          int j = v + d; if (j >= m) j -= m;
          const int lo = v < j ? v : j, hi = v < j ? j : v;
          mine |= pair_needs_rotation(M[Sh::addr(lo, hi)], dii, M[Sh::addr(j, j)]) ? 1 : 0;
        }
      }
      need = __syncthreads_or(mine) != 0;
      if (!need || sweeps >= max_num_sweeps) break;
      // pivots of the step inside the blocks, then  I | A B A B ... A B,  every step with the next step's pivots fused in
      if (rid >= 0 && rid < nbp) cta5_diag<T, NPB, EVECS, 2, false>(cx, rid, T(0), T(0), T(0), T(0), n_rot);
      __syncthreads();
      cta5_phase<T, NPB, EVECS, XW, 2, 0>(cx, e, n_rot);
      for (int bs = 0; bs < nb; bs += 2) {
        cta5_phase<T, NPB, EVECS, XW, 0, 1>(cx, e, n_rot);
        if (bs + 2 < nb) cta5_phase<T, NPB, EVECS, XW, 1, 0>(cx, e, n_rot);
        else cta5_phase<T, NPB, EVECS, XW, 1, -1>(cx, e, n_rot);
      }
      ++sweeps;
      if constexpr (EVECS) {
        // fold the row scales back into the rows (keeps f and 1/f near 1)
        if (worker) {
#pragma unroll
          for (int k = 0; k < 32; ++k) e[k] *= fr[32 * ch + k].x;
        }
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
      if (worker) {
#pragma unroll
        for (int k = 0; k < 32; ++k) Es[(32 * ch + k) * NPB + v] = e[k];
      }
      __syncthreads();
      for (int idx = tid; idx < n * n; idx += blockDim.x) {
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
