// jpd_cluster4.cuh -- tier 4 (129 <= n <= 256), second kernel: the block odd-even ordering of jpd_cta4.cuh
// on a 4-CTA thread-block cluster.
//
// Path being replaced: Jacobi::Diagonalize, /root/reference/include/jacobi_pd.hpp:159-220
// (CalcRot :233-256, ApplyRot :327-393, ApplyRotLeft :409-419, SortRows :477-508).
//
// The round-2 profile of jpd_cluster.cuh (one rotation step per trip of M through distributed shared
// memory, 3 cluster barriers per two steps) shows 2790 shared-memory wavefronts per step and CTA: like
// tiers 2 and 3 it is bound by shared-memory traffic.  This kernel is jpd_cta4.cuh spread over four SMs:
//  * block ordering: two rotation steps per block step, 4x4 tiles of M in registers, scaled
//    eigenvector rotations, two columns x 16 rows of E per thread (see jpd_cta4.cuh);
//  * M in DISTRIBUTED shared memory, tile-major: the 64 tile rows (4 matrix rows each) are dealt
//    to the CTAs in groups of 8, group g to CTA g or 7-g, which gives every CTA exactly a quarter of
//    the tiles of the upper triangle and makes a pattern-B tile -- it straddles two tile rows --
//    reach into a neighbour CTA only once in eight (ld/st.shared::cluster for its lower half);
//  * every CTA holds all (c,s), (alpha,beta) and row scales: the thread that computes a pivot writes
//    them into the four CTAs (st.shared::cluster), so every read is local;
//  * E in the registers of four SMs: CTA r owns positions 64r .. 64r+63; the rows next to a CTA
//    boundary are read from the neighbour's buffer (ld.shared::cluster);
//  * 2 cluster barriers per block step (= two rotation steps), plain release/acquire.
#pragma once
#include "jpd_cta4.cuh"
#include "jpd_cluster.cuh"

namespace jpd {

struct Cl4Shape {
  static constexpr int kN = 256;
  static constexpr int kCtas = 4;
  static constexpr int kThreads = 512;
  static constexpr int kBP = 64;                        // block pairs at n = 256
  static constexpr int kW = 65;                         // tiles per stored tile row (== 1 mod 16)
  static constexpr int kTileStride = 17;
  static constexpr int kCs = 5;
  static constexpr int kRowsLocal = 16;                 // tile rows per CTA
  static constexpr int kMElems = kRowsLocal * kW * kTileStride;   // 17680 >= 64 x 256 (eigenvector staging)
  static constexpr int kERows = 16, kEChunks = 4;       // E: 64 positions per CTA, 16 rows x 2 columns per thread
  // elements of T:  M | (c,s) | (alpha,beta) x2 | (f,1/f) | diagonal | boundary rows lo, hi
  static constexpr int kElems = kMElems + 2 * kCs * kBP + 2 * 2 * kCs * kBP + 2 * kN + kN + 2 * 2 * kEChunks * kN;
  static constexpr int kInts = kN + 16;                 // perm | flags[4] | rot[4] | pad
  JPD_CX static constexpr size_t table_bytes() { return (size_t)2 * kThreads * sizeof(unsigned); }
  JPD_CX static constexpr size_t block_bytes(size_t elem) {
    return table_bytes() + ((size_t)kElems * elem + 15) / 16 * 16 + (size_t)kInts * sizeof(int);
  }
  JPD_HD static int owner_row(int P) { const int g = P >> 3; return g < 4 ? g : 7 - g; }
  JPD_HD static int local_row(int P) { return ((P >> 3) < 4 ? 0 : 8) + (P & 7); }
  JPD_HD static int global_row(int l, int rank) { return l < 8 ? 8 * rank + l : 8 * (7 - rank) + (l - 8); }
  JPD_HD static int toff(int P, int Q) { return (local_row(P) * kW + Q) * kTileStride; }   // offset inside the owner's M
};

#if defined(__CUDACC__)

__device__ __forceinline__ void st_cluster_v2(unsigned a, double x, double y) {
  asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" :: "r"(a), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ void st_cluster_v2(unsigned a, float x, float y) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" :: "r"(a), "f"(x), "f"(y) : "memory");
}

template <typename T> struct Cl4Ctx {
  using T2 = typename Pair2<T>::type;
  T* M; T2* cs; T2* ab; T2* fr; T* dg; T* xlo; T* xhi;
  const unsigned* tblA; const unsigned* tblB;
  unsigned m_u32, cs_u32, ab_u32, fr_u32, dg_u32, xlo_u32, xhi_u32;   // this CTA's shared-window addresses
  int m, nbp, rank, c0, ech;
};

// write one (x,y) pair at the same offset of all four CTAs' copies of an array
template <typename T>
__device__ __forceinline__ void cl4_publish(unsigned base_u32, int index, T x, T y) {
  const unsigned a = base_u32 + (unsigned)(index * 2 * (int)sizeof(T));
#pragma unroll
  for (unsigned r = 0; r < 4; ++r) st_cluster_v2(cluster_map(a, r), x, y);
}

// ---- pivots of block pair J (owned by this CTA: its first tile row is here) in pattern MODE
template <typename T, bool EVECS, int MODE>
__device__ __forceinline__ void cl4_diag(const Cl4Ctx<T>& cx, int J, int& n_rot) {
  using Sh = Cl4Shape;
  using T2 = typename Pair2<T>::type;
  constexpr int OFF = (MODE == 1) ? 2 : 0;
  constexpr int W = Sh::kW;
  constexpr int esz = (int)sizeof(T);
  T* D = cx.M + Sh::toff(J, J);
  // pattern B: the entries with both indices in the second block (d2, d3, o23) live in tile (J+1, J+1), which
  // is in another CTA when J+1 starts a group of tile rows that this CTA does not own
  const bool remote = (MODE == 1) && Sh::owner_row(J + 1) != cx.rank;
  const unsigned rb = remote ? cluster_map(cx.m_u32 + (unsigned)(Sh::toff(J + 1, J + 1) * esz), (unsigned)Sh::owner_row(J + 1)) : 0u;
  T d0 = D[tile4_off<OFF, W>(0, 0)], d1 = D[tile4_off<OFF, W>(1, 1)];
  T o01 = D[tile4_off<OFF, W>(0, 1)], o02 = D[tile4_off<OFF, W>(0, 2)], o03 = D[tile4_off<OFF, W>(0, 3)];
  T o12 = D[tile4_off<OFF, W>(1, 2)], o13 = D[tile4_off<OFF, W>(1, 3)];
  T d2, d3, o23;
  if (remote) {   // local entries (0,0), (1,1), (0,1) of the remote tile
    d2 = ld_cluster(rb, T(0)); d3 = ld_cluster(rb + 5 * esz, T(0)); o23 = ld_cluster(rb + 1 * esz, T(0));
  } else {
    d2 = D[tile4_off<OFF, W>(2, 2)]; d3 = D[tile4_off<OFF, W>(3, 3)]; o23 = D[tile4_off<OFF, W>(2, 3)];
  }
  const int p0 = 4 * J + OFF;
  T f0 = T(1), f1 = T(1), f2 = T(1), f3 = T(1), r0 = T(1), r1 = T(1), r2 = T(1), r3 = T(1);
  if (EVECS) {   // (f, 1/f) of the four positions (every CTA holds all of them)
    const T2 s0 = cx.fr[p0], s1 = cx.fr[p0 + 1], s2 = cx.fr[p0 + 2], s3 = cx.fr[p0 + 3];
    f0 = s0.x; r0 = s0.y; f1 = s1.x; r1 = s1.y; f2 = s2.x; r2 = s2.y; f3 = s3.x; r3 = s3.y;
  }
  const int ci = Sh::kCs * J;
  const unsigned abu = cx.ab_u32 + (unsigned)(ab_buffer(MODE) * Sh::kCs * Sh::kBP * 2 * esz);
  T ca, sa, ta, ua, cb, sb, tb, ub, al, be;
  if (MODE != 2) {
    n_rot += pivot4<T, true>(d0, d2, o02, ca, sa, ta, ua) ? 1 : 0;
    n_rot += pivot4<T, true>(d1, d3, o13, cb, sb, tb, ub) ? 1 : 0;
    rotate_block4<T, true>(o01, o03, o12, o23, ca, sa, cb, sb);       // rows {0,2} x columns {1,3}
    cl4_publish<T>(cx.cs_u32, ci + 0, ca, sa);
    cl4_publish<T>(cx.cs_u32, ci + 1, cb, sb);
    if (EVECS) {
      scale_update4<T, true>(f0, f2, r0, r2, ca, ta, ua, al, be); cl4_publish<T>(abu, ci + 0, al, be);
      scale_update4<T, true>(f1, f3, r1, r3, cb, tb, ub, al, be); cl4_publish<T>(abu, ci + 1, al, be);
    }
    n_rot += pivot4<T, false>(d0, d3, o03, ca, sa, ta, ua) ? 1 : 0;
    n_rot += pivot4<T, false>(d1, d2, o12, cb, sb, tb, ub) ? 1 : 0;
    rotate_block4<T, false>(o01, o02, o13, o23, ca, sa, cb, sb);      // rows {0,3} x columns {1,2}
    cl4_publish<T>(cx.cs_u32, ci + 2, ca, sa);
    cl4_publish<T>(cx.cs_u32, ci + 3, cb, sb);
    if (EVECS) {
      scale_update4<T, false>(f0, f3, r0, r3, ca, ta, ua, al, be); cl4_publish<T>(abu, ci + 2, al, be);
      scale_update4<T, false>(f1, f2, r1, r2, cb, tb, ub, al, be); cl4_publish<T>(abu, ci + 3, al, be);
    }
  } else {
    n_rot += pivot4<T, false>(d0, d1, o01, ca, sa, ta, ua) ? 1 : 0;
    n_rot += pivot4<T, false>(d2, d3, o23, cb, sb, tb, ub) ? 1 : 0;
    rotate_block4<T, false>(o02, o03, o12, o13, ca, sa, cb, sb);      // rows {0,1} x columns {2,3}
    cl4_publish<T>(cx.cs_u32, ci + 0, ca, sa);
    cl4_publish<T>(cx.cs_u32, ci + 1, cb, sb);
    if (EVECS) {
      scale_update4<T, false>(f0, f1, r0, r1, ca, ta, ua, al, be); cl4_publish<T>(abu, ci + 0, al, be);
      scale_update4<T, false>(f2, f3, r2, r3, cb, tb, ub, al, be); cl4_publish<T>(abu, ci + 1, al, be);
    }
  }
  D[tile4_off<OFF, W>(0, 0)] = d0; D[tile4_off<OFF, W>(1, 1)] = d1;
  D[tile4_off<OFF, W>(0, 1)] = o01; D[tile4_off<OFF, W>(0, 2)] = o02; D[tile4_off<OFF, W>(0, 3)] = o03;
  D[tile4_off<OFF, W>(1, 2)] = o12; D[tile4_off<OFF, W>(1, 3)] = o13;
  if (remote) {
    st_cluster(rb, d2); st_cluster(rb + 5 * esz, d3); st_cluster(rb + 1 * esz, o23);
  } else {
    D[tile4_off<OFF, W>(2, 2)] = d2; D[tile4_off<OFF, W>(3, 3)] = d3; D[tile4_off<OFF, W>(2, 3)] = o23;
  }
  if (EVECS) {
    cl4_publish<T>(cx.fr_u32, p0, f0, r0);
    cl4_publish<T>(cx.fr_u32, p0 + 1, f1, r1);
    cl4_publish<T>(cx.fr_u32, p0 + 2, f2, r2);
    cl4_publish<T>(cx.fr_u32, p0 + 3, f3, r3);
  }
}

// ---- off-diagonal tiles owned by this CTA, and in pattern B its strips
template <typename T, int MODE>
__device__ __forceinline__ bool cl4_tiles(const Cl4Ctx<T>& cx) {     // returns: this thread wrote another CTA's memory
  using Sh = Cl4Shape;
  bool wrote_remote = false;
  using T2 = typename Pair2<T>::type;
  constexpr int OFF = (MODE == 1) ? 2 : 0;
  constexpr int W = Sh::kW;
  constexpr int esz = (int)sizeof(T);
  const int npairs = (MODE == 1) ? cx.nbp - 1 : cx.nbp;
  const int tid = threadIdx.x;
  const unsigned ent = (MODE == 1 ? cx.tblB : cx.tblA)[tid];
  if (ent != 0xffffffffu) {
    const int P = (ent >> 16) & 0x3f, Q = (ent >> 22) & 0x3f;
    T* B = cx.M + (ent & 0xffffu);
    const bool remote = (MODE == 1) && (ent >> 28) != 0;
    wrote_remote = remote;
    const unsigned rb = remote ? cluster_map(cx.m_u32 + (unsigned)(Sh::toff(P + 1, Q) * esz), (unsigned)Sh::owner_row(P + 1)) : 0u;
    T t[16];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (MODE == 1 && r >= 2 && remote)   // the lower half lives in the neighbour: offsets relative to ITS tile (P+1, Q)
          t[4 * r + c] = ld_cluster(rb + (unsigned)((tile4_off<2, W>(r, c) - W * Sh::kTileStride) * esz), T(0));
        else
          t[4 * r + c] = B[tile4_off<OFF, W>(r, c)];
      }
    const T2* csP = cx.cs + Sh::kCs * P;
    const T2* csQ = cx.cs + Sh::kCs * Q;
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
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (MODE == 1 && r >= 2 && remote)
          st_cluster(rb + (unsigned)((tile4_off<2, W>(r, c) - W * Sh::kTileStride) * esz), t[4 * r + c]);
        else
          B[tile4_off<OFF, W>(r, c)] = t[4 * r + c];
      }
  }
  // ---- pattern B: rows {0,1} x column block pair j (CTA 0, threads from the end), and row block pair j x
  // columns {m-2, m-1} (the CTA that owns tile row j; its lower two rows may live in the neighbour)
  if (MODE == 1) {
    const int r = Sh::kThreads - 1 - tid;
    if (r < 64) {
      const int j = r;
      if (cx.rank == 0 && j < npairs) {      // top strip j: tile row 0, tiles (0,j) columns 2,3 and (0,j+1) columns 0,1
        T* S = cx.M + Sh::toff(0, j);
        T* s0 = S + 2; T* s1 = S + 3; T* s2 = S + Sh::kTileStride; T* s3 = S + Sh::kTileStride + 1;
        T a0 = s0[0], a1 = s1[0], a2 = s2[0], a3 = s3[0];
        T b0 = s0[4], b1 = s1[4], b2 = s2[4], b3 = s3[4];
        const T2* g = cx.cs + Sh::kCs * j;
        const T2 g0 = g[0], g1 = g[1], g2 = g[2], g3 = g[3];
        rotate_pair4<T, true>(a0, a2, g0.x, g0.y); rotate_pair4<T, true>(b0, b2, g0.x, g0.y);
        rotate_pair4<T, true>(a1, a3, g1.x, g1.y); rotate_pair4<T, true>(b1, b3, g1.x, g1.y);
        rotate_pair4<T, false>(a0, a3, g2.x, g2.y); rotate_pair4<T, false>(b0, b3, g2.x, g2.y);
        rotate_pair4<T, false>(a1, a2, g3.x, g3.y); rotate_pair4<T, false>(b1, b2, g3.x, g3.y);
        s0[0] = a0; s1[0] = a1; s2[0] = a2; s3[0] = a3;
        s0[4] = b0; s1[4] = b1; s2[4] = b2; s3[4] = b3;
      }
    } else if (r < 64 + 16) {
      const int j = Sh::global_row(r - 64, cx.rank);
      if (j < npairs) {                      // right strip j: rows 4j+2..4j+5 x columns m-2, m-1 (tile column nbp-1, local columns 2,3)
        const int Qc = cx.nbp - 1;
        T* S = cx.M + Sh::toff(j, Qc) + 2;   // tile (j, Qc), column 2
        const bool remote = Sh::owner_row(j + 1) != cx.rank;
        wrote_remote = wrote_remote || remote;
        const unsigned rb = remote ? cluster_map(cx.m_u32 + (unsigned)((Sh::toff(j + 1, Qc) + 2) * esz), (unsigned)Sh::owner_row(j + 1)) : 0u;
        // k = 0,1: rows 2,3 of tile (j, Qc); k = 2,3: rows 0,1 of tile (j+1, Qc)
        T a0 = S[8], a1 = S[12], b0 = S[9], b1 = S[13];
        T a2, a3, b2, b3;
        if (remote) { a2 = ld_cluster(rb, T(0)); a3 = ld_cluster(rb + 4 * esz, T(0)); b2 = ld_cluster(rb + esz, T(0)); b3 = ld_cluster(rb + 5 * esz, T(0)); }
        else { T* S2 = S + W * Sh::kTileStride; a2 = S2[0]; a3 = S2[4]; b2 = S2[1]; b3 = S2[5]; }
        const T2* g = cx.cs + Sh::kCs * j;
        const T2 g0 = g[0], g1 = g[1], g2 = g[2], g3 = g[3];
        rotate_pair4<T, true>(a0, a2, g0.x, g0.y); rotate_pair4<T, true>(b0, b2, g0.x, g0.y);
        rotate_pair4<T, true>(a1, a3, g1.x, g1.y); rotate_pair4<T, true>(b1, b3, g1.x, g1.y);
        rotate_pair4<T, false>(a0, a3, g2.x, g2.y); rotate_pair4<T, false>(b0, b3, g2.x, g2.y);
        rotate_pair4<T, false>(a1, a2, g3.x, g3.y); rotate_pair4<T, false>(b1, b2, g3.x, g3.y);
        S[8] = a0; S[12] = a1; S[9] = b0; S[13] = b1;
        if (remote) { st_cluster(rb, a2); st_cluster(rb + 4 * esz, a3); st_cluster(rb + esz, b2); st_cluster(rb + 5 * esz, b3); }
        else { T* S2 = S + W * Sh::kTileStride; S2[0] = a2; S2[4] = a3; S2[1] = b2; S2[5] = b3; }
      }
    }
  }
  return wrote_remote;
}

// ---- eigenvector rows of a block step of mode MODE (see cta4_evecs); the rows next to a CTA boundary come
// from the neighbour CTA's buffer
template <typename T, int MODE>
__device__ __forceinline__ void cl4_evecs(const Cl4Ctx<T>& cx, T (&e)[32]) {
  using Sh = Cl4Shape;
  using T2 = typename Pair2<T>::type;
  constexpr int OFF = (MODE == 1) ? 2 : 0;
  constexpr int esz = (int)sizeof(T);
  const int npairs = (MODE == 1) ? cx.nbp - 1 : cx.nbp;
  const int ech = cx.ech;
  const T2* ab = cx.ab + ab_buffer(MODE) * Sh::kCs * Sh::kBP;
  const int J0 = 16 * cx.rank + 4 * ech;    // block pair of this chunk's first four rows
  if (MODE == 1) {
    if ((ech > 0 || cx.rank > 0) && J0 - 1 < npairs) {       // the block pair shared with the chunk below
      const T2* h = ab + Sh::kCs * (J0 - 1);
      const T2 h0 = h[0], h1 = h[1], h2 = h[2], h3 = h[3];
#pragma unroll
      for (int col = 0; col < 2; ++col) {
        const int c = cx.c0 + col * (Sh::kN / 2), cb = Sh::kERows * col;
        T y0, y1;
        if (ech > 0) { y0 = cx.xhi[(2 * (ech - 1) + 0) * Sh::kN + c]; y1 = cx.xhi[(2 * (ech - 1) + 1) * Sh::kN + c]; }
        else {
          const unsigned a = cluster_map(cx.xhi_u32 + (unsigned)(((2 * 3 + 0) * Sh::kN + c) * esz), (unsigned)(cx.rank - 1));
          y0 = ld_cluster(a, T(0)); y1 = ld_cluster(a + (unsigned)(Sh::kN * esz), T(0));
        }
        T lo0 = e[cb], lo1 = e[cb + 1];
        fast_rotate_group<T, T2, MODE>(y0, y1, lo0, lo1, h0, h1, h2, h3);
        e[cb] = lo0; e[cb + 1] = lo1;
      }
    }
    if ((ech < 3 || cx.rank < 3) && J0 + 3 < npairs) {       // the block pair shared with the chunk above
      const T2* h = ab + Sh::kCs * (J0 + 3);
      const T2 h0 = h[0], h1 = h[1], h2 = h[2], h3 = h[3];
#pragma unroll
      for (int col = 0; col < 2; ++col) {
        const int c = cx.c0 + col * (Sh::kN / 2), cb = Sh::kERows * col;
        T y2, y3;
        if (ech < 3) { y2 = cx.xlo[(2 * (ech + 1) + 0) * Sh::kN + c]; y3 = cx.xlo[(2 * (ech + 1) + 1) * Sh::kN + c]; }
        else {
          const unsigned a = cluster_map(cx.xlo_u32 + (unsigned)(((2 * 0 + 0) * Sh::kN + c) * esz), (unsigned)(cx.rank + 1));
          y2 = ld_cluster(a, T(0)); y3 = ld_cluster(a + (unsigned)(Sh::kN * esz), T(0));
        }
        T hi0 = e[cb + 14], hi1 = e[cb + 15];
        fast_rotate_group<T, T2, MODE>(hi0, hi1, y2, y3, h0, h1, h2, h3);
        e[cb + 14] = hi0; e[cb + 15] = hi1;
      }
    }
  }
#pragma unroll
  for (int jj = 0; jj < 4 - (MODE == 1 ? 1 : 0); ++jj) {
    if (J0 + jj < npairs) {
      const T2* h = ab + Sh::kCs * (J0 + jj);
      const T2 h0 = h[0], h1 = h[1], h2 = (MODE != 2) ? h[2] : h[0], h3 = (MODE != 2) ? h[3] : h[1];
      const int p = 4 * jj + OFF;
#pragma unroll
      for (int col = 0; col < 2; ++col) {
        const int q = Sh::kERows * col + p;
        fast_rotate_group<T, T2, MODE>(e[q], e[q + 1], e[q + 2], e[q + 3], h0, h1, h2, h3);
      }
    }
  }
}

// first phase of a block step: the pivots this CTA owns (16 threads) beside the eigenvector phase of the previous step
template <typename T, bool EVECS, int MODE, int PREV>
__device__ __forceinline__ bool cl4_front(const Cl4Ctx<T>& cx, T (&e)[EVECS ? 32 : 1], int& n_rot) {   // returns: wrote another CTA's memory
  using Sh = Cl4Shape;
  const int npairs = (MODE == 1) ? cx.nbp - 1 : cx.nbp;
  bool wrote_remote = false;
  if (threadIdx.x < 16) {
    const int J = Sh::global_row(threadIdx.x, cx.rank);
    if (J < npairs) { cl4_diag<T, EVECS, MODE>(cx, J, n_rot); wrote_remote = true; }
  }
  if constexpr (EVECS) {
    if constexpr (PREV >= 0) cl4_evecs<T, PREV>(cx, e);
    if constexpr (MODE == 1) {   // the rows next to the chunk boundaries, for this step's eigenvector phase
#pragma unroll
      for (int col = 0; col < 2; ++col) {
        const int c = cx.c0 + col * (Sh::kN / 2), cb = Sh::kERows * col;
        cx.xlo[(2 * cx.ech + 0) * Sh::kN + c] = e[cb]; cx.xlo[(2 * cx.ech + 1) * Sh::kN + c] = e[cb + 1];
        cx.xhi[(2 * cx.ech + 0) * Sh::kN + c] = e[cb + 14]; cx.xhi[(2 * cx.ech + 1) * Sh::kN + c] = e[cb + 15];
      }
    }
  }
  return wrote_remote;
}

// The two barriers of a block step are plain barrier.cluster arrive.release / wait.acquire.  The relaxed arrive
// with the release spelled out as fence.acq_rel.cluster in the threads that wrote remotely plus one cumulative
// fence per CTA (the scheme of jpd_cluster.cuh) was measured too: 95.7 ms against 94.0 ms for 528 matrices of
// 256 x 256 -- here a sixteenth of the threads of EVERY warp write into a neighbour in the tile phase, so every warp
// pays its fence anyway.
__device__ __forceinline__ void cl4_sync_after_pivots(bool) { cluster_barrier(); }
__device__ __forceinline__ void cl4_sync_after_tiles(bool) { cluster_barrier(); }

template <typename T, bool EVECS>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(512, 1)
jpd_cluster4_kernel(const T* __restrict__ mat, T* __restrict__ evals, T* __restrict__ evecs,
                    int* __restrict__ n_iters, int n, long long batch, int soa, int sort_criteria,
                    int max_num_sweeps) {
  using Sh = Cl4Shape;
  using T2 = typename Pair2<T>::type;
  constexpr int esz = (int)sizeof(T);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int rank = (int)cluster_ctarank();
  const int m = (n + 3) & ~3;
  const int nbp = m >> 2, nb = m >> 1;

  unsigned* tblA = reinterpret_cast<unsigned*>(smem_raw);
  unsigned* tblB = tblA + Sh::kThreads;
  T* M = reinterpret_cast<T*>(smem_raw + Sh::table_bytes());
  T2* cs = reinterpret_cast<T2*>(M + Sh::kMElems);
  T2* ab = cs + Sh::kCs * Sh::kBP;
  T2* fr = ab + 2 * Sh::kCs * Sh::kBP;
  T* dg = reinterpret_cast<T*>(fr + Sh::kN);
  T* xlo = dg + Sh::kN;
  T* xhi = xlo + 2 * Sh::kEChunks * Sh::kN;
  int* perm = reinterpret_cast<int*>(smem_raw + Sh::table_bytes() + ((size_t)Sh::kElems * sizeof(T) + 15) / 16 * 16);
  int* flags = perm + Sh::kN;      // [4]
  int* rots = flags + 4;           // [4]
  int* ovf = rots + 4;             // scratch of the table builder: [2][count]  (tables are final before use)

  Cl4Ctx<T> cx;
  cx.M = M; cx.cs = cs; cx.ab = ab; cx.fr = fr; cx.dg = dg; cx.xlo = xlo; cx.xhi = xhi;
  cx.tblA = tblA; cx.tblB = tblB;
  cx.m_u32 = smem_u32(M); cx.cs_u32 = smem_u32(cs); cx.ab_u32 = smem_u32(ab); cx.fr_u32 = smem_u32(fr);
  cx.dg_u32 = smem_u32(dg); cx.xlo_u32 = smem_u32(xlo); cx.xhi_u32 = smem_u32(xhi);
  cx.m = m; cx.nbp = nbp; cx.rank = rank; cx.c0 = tid & 127; cx.ech = tid >> 7;

  // ---- tile tables: the off-diagonal tiles whose tile row this CTA owns, one per thread; the owner thread of a
  // tile has tid % 16 == bank class (local row + Q) mod 16 where the class has room, any free thread otherwise
  for (int i = tid; i < 2 * Sh::kThreads; i += Sh::kThreads) tblA[i] = 0xffffffffu;
  if (tid < 2) ovf[tid] = 0;
  __syncthreads();
  if (tid < 32) {
    const int cl = tid & 15, pat = tid >> 4;
    unsigned* tbl = pat == 0 ? tblA : tblB;
    const int np = pat == 0 ? nbp : nbp - 1;
    int slot = 0;
    for (int l = 0; l < Sh::kRowsLocal; ++l) {
      const int P = Sh::global_row(l, rank);
      if (P >= np) continue;
      for (int Q = P + 1 + ((cl - l - P - 1) & 15); Q < np; Q += 16) {     // (l + Q) mod 16 == cl
        const unsigned remote = (pat == 1 && Sh::owner_row(P + 1) != rank) ? 1u : 0u;
        const unsigned ent = (unsigned)((l * Sh::kW + Q) * Sh::kTileStride) | ((unsigned)P << 16) | ((unsigned)Q << 22) | (remote << 28);
        if (slot < Sh::kThreads / 16) { tbl[slot * 16 + cl] = ent; ++slot; }
        else {   // class over-full: kept in a per-pattern list in M's storage (M is not loaded yet) for the second pass
          const int k = atomicAdd(&ovf[pat], 1);
          reinterpret_cast<unsigned*>(M)[pat * 256 + k] = ent;
        }
      }
    }
  }
  __syncthreads();
  if (tid < 2) {
    unsigned* tbl = tid == 0 ? tblA : tblB;
    const int cnt = ovf[tid];
    int free_slot = 0;
    for (int k = 0; k < cnt; ++k) {
      while (free_slot < Sh::kThreads && tbl[free_slot] != 0xffffffffu) ++free_slot;
      if (free_slot < Sh::kThreads) tbl[free_slot++] = reinterpret_cast<unsigned*>(M)[tid * 256 + k];
    }
  }
  cluster_barrier();

  const long long n_clusters = gridDim.x / Sh::kCtas;
  for (long long b = blockIdx.x / Sh::kCtas; b < batch; b += n_clusters) {
    // ---- load the tile rows this CTA owns: upper triangle only (jacobi_pd.hpp:167-171)
    for (int i = tid; i < Sh::kMElems; i += Sh::kThreads) M[i] = T(0);
    if (tid < Sh::kN) { T2 one; one.x = T(1); one.y = T(1); fr[tid] = one; }
    __syncthreads();
    for (int idx = tid; idx < 4 * Sh::kRowsLocal * n; idx += Sh::kThreads) {
      const int lr = idx / n, c = idx - lr * n;          // lr: local matrix row = 4 * local tile row + row inside the tile
      const int r = 4 * Sh::global_row(lr >> 2, rank) + (lr & 3);
      if (r < n && c >= r) {
        const size_t e_idx = (size_t)r * n + c;
        const T val = soa ? mat[e_idx * (size_t)batch + (size_t)b] : mat[(size_t)b * n * n + e_idx];
        M[((lr >> 2) * Sh::kW + (c >> 2)) * Sh::kTileStride + 4 * (lr & 3) + (c & 3)] = val;
      }
    }
    T e[EVECS ? 32 : 1];
    if (EVECS) {
#pragma unroll
      for (int k = 0; k < 32; ++k)   // :172-178
        e[k] = (64 * rank + Sh::kERows * cx.ech + (k & 15) == cx.c0 + (k >> 4) * (Sh::kN / 2)) ? T(1) : T(0);
    }
    cluster_barrier();

    // ---- sweeps
    int n_rot = 0, sweeps = 0;
    bool need = false;
    for (;;) {
      // every CTA gets a copy of the diagonal ...
      if (tid < 4 * Sh::kRowsLocal) {
        const int l = tid >> 2, k = tid & 3;
        const int i = 4 * Sh::global_row(l, rank) + k;
        if (i < m) {
          const T d = M[(l * Sh::kW + Sh::global_row(l, rank)) * Sh::kTileStride + 5 * k];
#pragma unroll
          for (unsigned r = 0; r < 4; ++r) st_cluster(cluster_map(cx.dg_u32 + (unsigned)(i * esz), r), d);
        }
      }
      cluster_barrier();
      // ... and tests the entries of its own tiles with the reference's rule (:191)
      int mine = 0;
      for (int idx = tid; idx < Sh::kRowsLocal * nbp; idx += Sh::kThreads) {
        const int l = idx / nbp, Q = idx - l * nbp;
        const int P = Sh::global_row(l, rank);
        if (P < nbp && Q >= P) {
          const T* B = M + (l * Sh::kW + Q) * Sh::kTileStride;
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int i = 4 * P + r, j = 4 * Q + c;
              if (j > i) mine |= pair_needs_rotation(B[4 * r + c], dg[i], dg[j]) ? 1 : 0;
            }
        }
      }
      const int cta_need = __syncthreads_or(mine);
      if (tid < 4) st_cluster_i32(cluster_map(smem_u32(flags) + (unsigned)(rank * 4), (unsigned)tid), cta_need);
      cluster_barrier();
      need = (flags[0] | flags[1] | flags[2] | flags[3]) != 0;
      if (!need || sweeps >= max_num_sweeps) break;

      cl4_sync_after_pivots(cl4_front<T, EVECS, 2, -1>(cx, e, n_rot));
      cl4_sync_after_tiles(cl4_tiles<T, 2>(cx));
      for (int bs = 0; bs < nb; bs += 2) {
        if (bs == 0) cl4_sync_after_pivots(cl4_front<T, EVECS, 0, 2>(cx, e, n_rot));
        else cl4_sync_after_pivots(cl4_front<T, EVECS, 0, 1>(cx, e, n_rot));
        cl4_sync_after_tiles(cl4_tiles<T, 0>(cx));
        cl4_sync_after_pivots(cl4_front<T, EVECS, 1, 0>(cx, e, n_rot));
        cl4_sync_after_tiles(cl4_tiles<T, 1>(cx));
      }
      ++sweeps;
      if constexpr (EVECS) {
        cl4_evecs<T, 1>(cx, e);     // the last B step's eigenvector phase
#pragma unroll
        for (int k = 0; k < Sh::kERows; ++k) {   // fold the row scales back into the rows
          const T fk = fr[64 * rank + Sh::kERows * cx.ech + k].x;
          e[k] *= fk; e[Sh::kERows + k] *= fk;
        }
        cluster_barrier();   // every CTA is done with its neighbours' boundary rows and with the scales
        if (tid < Sh::kN) { T2 one; one.x = T(1); one.y = T(1); fr[tid] = one; }
      }
    }
    const bool converged = !need && max_num_sweeps > 0;
    // rotations: block sum, then every CTA tells CTA 0
    {
      __shared__ int block_rot;
      if (tid == 0) block_rot = 0;
      __syncthreads();
      if (n_rot) atomicAdd(&block_rot, n_rot);
      __syncthreads();
      if (tid == 0) st_cluster_i32(cluster_map(smem_u32(rots) + (unsigned)(rank * 4), 0u), block_rot);
    }

    // ---- eigenvalues (:207-209) from the copy of the diagonal.  A sweep reverses the order of the BLOCKS:
    // position p holds original index p after an even number of sweeps, 2 (nb-1 - p/2) + (p&1) after an odd one.
    auto orig_index = [&](int p) { return (sweeps & 1) ? 2 * (nb - 1 - (p >> 1)) + (p & 1) : p; };
    const bool real = tid < m && orig_index(tid) < n;
    const T my_ev = tid < m ? dg[tid] : T(0);
    if (real) {
      int rk = 0;
      const bool sorted = sort_criteria >= SORT_DECREASING_EVALS && sort_criteria <= SORT_INCREASING_ABS_EVALS;
      const bool use_abs = sort_criteria >= SORT_DECREASING_ABS_EVALS;
      const bool decreasing = (sort_criteria == SORT_DECREASING_EVALS) || (sort_criteria == SORT_DECREASING_ABS_EVALS);
      const long long key = sortable_bits(use_abs ? jabs(my_ev) : my_ev);
      for (int j = 0; j < m; ++j) {
        if (orig_index(j) >= n) continue;
        bool before = j < tid;
        if (sorted) {
          const T o = dg[j];
          const long long other = sortable_bits(use_abs ? jabs(o) : o);
          before = (decreasing ? (other > key) : (other < key)) || (other == key && j < tid);
        }
        rk += before ? 1 : 0;
      }
      perm[rk] = tid;
      if (rank == 0) {
        if (soa) evals[(size_t)rk * (size_t)batch + (size_t)b] = my_ev;
        else evals[(size_t)b * n + rk] = my_ev;
      }
    }
    cluster_barrier();     // perm complete; every CTA's rotation count has arrived in CTA 0
    if (rank == 0 && tid == 0 && n_iters) n_iters[b] = converged ? rots[0] + rots[1] + rots[2] + rots[3] + 1 : 0;
    // ---- eigenvectors: this CTA's 64 rows staged in M's storage, then the sorted rows it owns go out
    if (EVECS) {
      T* Es = M;
#pragma unroll
      for (int k = 0; k < 32; ++k) Es[(Sh::kERows * cx.ech + (k & 15)) * Sh::kN + cx.c0 + (k >> 4) * (Sh::kN / 2)] = e[k];
      __syncthreads();
      for (int idx = tid; idx < n * n; idx += Sh::kThreads) {
        const int k = idx / n, c = idx - k * n;
        const int pos = perm[k] - 64 * rank;
        if (pos >= 0 && pos < 64) {
          const T val = Es[pos * Sh::kN + c];
          if (soa) evecs[(size_t)idx * (size_t)batch + (size_t)b] = val;
          else evecs[(size_t)b * n * n + idx] = val;
        }
      }
    }
    cluster_barrier();   // M / dg / perm / counters are reused by the next matrix of this cluster
  }
}

#endif  // __CUDACC__

}  // namespace jpd
