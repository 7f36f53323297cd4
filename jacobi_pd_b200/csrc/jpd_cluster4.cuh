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
//  * every CTA holds the pivot records -- (c,s) x4 and (alpha,beta) x4 -- of all 64 block pairs, so every read
//    in the tile and eigenvector phases is local.  The 16 pivot threads of a CTA write their records into the
//    CTA's own copy; warp 0 then sends the CTA's two contiguous ranges of records to the three other CTAs as bulk
//    copies (cp.async.bulk shared::cta -> shared::cluster) that complete on the RECEIVER's mbarrier, and every
//    thread of the cluster waits on its own CTA's mbarrier instead of a cluster barrier.  The row scales (f, 1/f)
//    live with the owner of the position's tile row.  (Measured with clock64, tests/scratch/t4prof.py: the 48 st.shared::cluster.v2 per pivot thread of
//    the first version cost 44 cycles each -- 2100 of the 2900 cycles of the pivot phase -- and the
//    MEMBAR.ALL.GPU of the release barrier behind them another 1100.)
//  * E in the registers of four SMs: CTA r owns positions 64r .. 64r+63; the rows next to a CTA boundary are
//    pushed into the neighbour's buffer (st.shared::cluster) after the pattern-A step;
//  * per block step (= two rotation steps): one mbarrier wait (pivots published) and one cluster barrier with
//    release/acquire (tiles and eigenvector rows done).
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
  static constexpr int kRec = 9;                        // per block pair: (c,s) x4 | (alpha,beta) x4 | pad (odd stride: banks)
  static constexpr int kRowsLocal = 16;                 // tile rows per CTA
  static constexpr int kMElems = kRowsLocal * kW * kTileStride;   // 17680 >= 64 x 256 (eigenvector staging)
  static constexpr int kERows = 16, kEChunks = 4;       // E: 64 positions per CTA, 16 rows x 2 columns per thread
  // elements of T:  M | pivot records | (f,1/f) | diagonal | boundary rows lo, hi
  static constexpr int kElems = kMElems + 2 * kRec * kBP + 2 * kN + kN + 2 * 2 * kEChunks * kN;
  static constexpr int kInts = kN + 16;                 // perm | flags[4] | rot[4] | scratch[2] | pad[2] | mbarrier (8 bytes)
  static constexpr int kBarInt = kN + 12;
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

// mbarrier / bulk-copy helpers (PTX ISA: mbarrier, cp.async.bulk)
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init_cluster() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {
  unsigned ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  unsigned spins = 0;
  while (!mbar_try_wait(bar, parity)) { if (++spins > (1u << 26)) __trap(); }   // a lost copy must not hang the device
}
// `bytes` (a multiple of 16) from this CTA's shared memory to a peer's, completing on the peer's mbarrier
__device__ __forceinline__ void bulk_copy_to_peer(unsigned dst_cluster, unsigned src_cta, unsigned bytes, unsigned bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(bar_cluster) : "memory");
}

// Everything in shared memory sits at a compile-time offset from M, so the context carries ONE pointer and one
// shared-window address (the kernel runs at the 128-register cap; a dozen loop-invariant pointers spilled).
template <typename T> struct Cl4Ctx {
  using T2 = typename Pair2<T>::type;
  using Sh = Cl4Shape;
  static constexpr int kIntsOff = (int)(((size_t)Sh::kElems * sizeof(T) + 15) / 16 * 16);   // bytes from M to the int region
  T* M;
  unsigned m_u32;                  // M's address in this CTA's shared window
  int m, nbp, rank, c0, ech;
  __device__ __forceinline__ T2* cs() const { return reinterpret_cast<T2*>(M + Sh::kMElems); }
  __device__ __forceinline__ T2* ab() const { return cs() + 4; }          // same records, second half
  __device__ __forceinline__ T2* fr() const { return cs() + Sh::kRec * Sh::kBP; }
  __device__ __forceinline__ T* dg() const { return reinterpret_cast<T*>(fr() + Sh::kN); }
  __device__ __forceinline__ T* xlo() const { return dg() + Sh::kN; }
  __device__ __forceinline__ T* xhi() const { return xlo() + 2 * Sh::kEChunks * Sh::kN; }
  __device__ __forceinline__ int* ints() const { return reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(M) + kIntsOff); }
  __device__ __forceinline__ int* rot_count() const { return ints() + Sh::kN + 10; }   // rotations of this CTA's pivots
  __device__ __forceinline__ const unsigned* tblA() const { return reinterpret_cast<const unsigned*>(reinterpret_cast<const unsigned char*>(M) - Sh::table_bytes()); }
  __device__ __forceinline__ const unsigned* tblB() const { return tblA() + Sh::kThreads; }
  __device__ __forceinline__ unsigned cs_u32() const { return m_u32 + (unsigned)(Sh::kMElems * sizeof(T)); }
  __device__ __forceinline__ unsigned fr_u32() const { return cs_u32() + (unsigned)(Sh::kRec * Sh::kBP * 2 * sizeof(T)); }
  __device__ __forceinline__ unsigned dg_u32() const { return fr_u32() + (unsigned)(Sh::kN * 2 * sizeof(T)); }
  __device__ __forceinline__ unsigned xlo_u32() const { return dg_u32() + (unsigned)(Sh::kN * sizeof(T)); }
  __device__ __forceinline__ unsigned xhi_u32() const { return xlo_u32() + (unsigned)(2 * Sh::kEChunks * Sh::kN * sizeof(T)); }
  __device__ __forceinline__ unsigned bar_u32() const { return m_u32 + (unsigned)kIntsOff + (unsigned)(Sh::kBarInt * sizeof(int)); }
};

// one (x,y) pair of this CTA's copy of an array (cl4_exchange_pivots sends it to the other CTAs)
template <typename T>
__device__ __forceinline__ void cl4_publish(typename Pair2<T>::type* base, int index, T x, T y) {
  typename Pair2<T>::type w; w.x = x; w.y = y;
  base[index] = w;
}

// ---- pivots of block pair J (owned by this CTA: its first tile row is here) in pattern MODE
template <typename T, bool EVECS, int MODE>
__device__ __forceinline__ void cl4_diag(const Cl4Ctx<T>& cx, int J) {
  int n_rot = 0;
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
  // (f, 1/f) of position p live in the CTA that owns tile row p / 4: the last two of a pattern-B pair whose second
  // block starts a neighbour's group of tile rows are remote, like d2, d3, o23
  const unsigned rfr = remote ? cluster_map(cx.fr_u32() + (unsigned)((p0 + 2) * 2 * esz), (unsigned)Sh::owner_row(J + 1)) : 0u;
  // the scales are loaded only when the first pivots are done (lds_after): 16 fewer live registers while they run
  auto load_scales = [&](T after) {
    const T2 s0 = lds_after(cx.fr() + p0, after), s1 = lds_after(cx.fr() + p0 + 1, after);
    f0 = s0.x; r0 = s0.y; f1 = s1.x; r1 = s1.y;
    if (remote) {
      f2 = ld_cluster(rfr, T(0)); r2 = ld_cluster(rfr + esz, T(0)); f3 = ld_cluster(rfr + 2 * esz, T(0)); r3 = ld_cluster(rfr + 3 * esz, T(0));
    } else {
      const T2 s2 = lds_after(cx.fr() + p0 + 2, after), s3 = lds_after(cx.fr() + p0 + 3, after);
      f2 = s2.x; r2 = s2.y; f3 = s3.x; r3 = s3.y;
    }
  };
  const int ci = Sh::kRec * J;
  T2* abu = cx.ab();
  T ca, sa, ta, ua, cb, sb, tb, ub, al, be;
  if (MODE != 2) {
    n_rot += pivot4<T, true>(d0, d2, o02, ca, sa, ta, ua) ? 1 : 0;
    n_rot += pivot4<T, true>(d1, d3, o13, cb, sb, tb, ub) ? 1 : 0;
    rotate_block4<T, true>(o01, o03, o12, o23, ca, sa, cb, sb);       // rows {0,2} x columns {1,3}
    cl4_publish<T>(cx.cs(), ci + 0, ca, sa);
    cl4_publish<T>(cx.cs(), ci + 1, cb, sb);
    if (EVECS) {
      load_scales(ca);
      scale_update4<T, true>(f0, f2, r0, r2, ca, ta, ua, al, be); cl4_publish<T>(abu, ci + 0, al, be);
      scale_update4<T, true>(f1, f3, r1, r3, cb, tb, ub, al, be); cl4_publish<T>(abu, ci + 1, al, be);
    }
    n_rot += pivot4<T, false>(d0, d3, o03, ca, sa, ta, ua) ? 1 : 0;
    n_rot += pivot4<T, false>(d1, d2, o12, cb, sb, tb, ub) ? 1 : 0;
    rotate_block4<T, false>(o01, o02, o13, o23, ca, sa, cb, sb);      // rows {0,3} x columns {1,2}
    cl4_publish<T>(cx.cs(), ci + 2, ca, sa);
    cl4_publish<T>(cx.cs(), ci + 3, cb, sb);
    if (EVECS) {
      scale_update4<T, false>(f0, f3, r0, r3, ca, ta, ua, al, be); cl4_publish<T>(abu, ci + 2, al, be);
      scale_update4<T, false>(f1, f2, r1, r2, cb, tb, ub, al, be); cl4_publish<T>(abu, ci + 3, al, be);
    }
  } else {
    n_rot += pivot4<T, false>(d0, d1, o01, ca, sa, ta, ua) ? 1 : 0;
    n_rot += pivot4<T, false>(d2, d3, o23, cb, sb, tb, ub) ? 1 : 0;
    rotate_block4<T, false>(o02, o03, o12, o13, ca, sa, cb, sb);      // rows {0,1} x columns {2,3}
    cl4_publish<T>(cx.cs(), ci + 0, ca, sa);
    cl4_publish<T>(cx.cs(), ci + 1, cb, sb);
    if (EVECS) {
      load_scales(ca);
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
    cl4_publish<T>(cx.fr(), p0, f0, r0);
    cl4_publish<T>(cx.fr(), p0 + 1, f1, r1);
    if (remote) {
      st_cluster(rfr, f2); st_cluster(rfr + esz, r2); st_cluster(rfr + 2 * esz, f3); st_cluster(rfr + 3 * esz, r3);
    } else {
      cl4_publish<T>(cx.fr(), p0 + 2, f2, r2);
      cl4_publish<T>(cx.fr(), p0 + 3, f3, r3);
    }
  }
  if (n_rot) atomicAdd(cx.rot_count(), n_rot);
}

// ---- off-diagonal tiles owned by this CTA, and in pattern B its strips
template <typename T, int MODE>
__device__ __forceinline__ void cl4_tiles(const Cl4Ctx<T>& cx) {
  using Sh = Cl4Shape;
  using T2 = typename Pair2<T>::type;
  constexpr int OFF = (MODE == 1) ? 2 : 0;
  constexpr int W = Sh::kW;
  constexpr int esz = (int)sizeof(T);
  const int npairs = (MODE == 1) ? cx.nbp - 1 : cx.nbp;
  const int tid = threadIdx.x;
  const unsigned ent = (MODE == 1 ? cx.tblB() : cx.tblA())[tid];
  if (ent != 0xffffffffu) {
    const int P = (ent >> 16) & 0x3f, Q = (ent >> 22) & 0x3f;
    T* B = cx.M + (ent & 0xffffu);
    const bool remote = (MODE == 1) && (ent >> 28) != 0;
    const unsigned rb = remote ? cluster_map(cx.m_u32 + (unsigned)(Sh::toff(P + 1, Q) * esz), (unsigned)Sh::owner_row(P + 1)) : 0u;
    T t[16];
    const T2* csP = cx.cs() + Sh::kRec * P;
    const T2* csQ = cx.cs() + Sh::kRec * Q;
    // element (r,c) of the tile; for r >= 2 of a straddling pattern-B tile the lower half lives in the neighbour
    // (offsets relative to ITS tile (P+1, Q)).  AFTER: see lds_after.
    auto elem = [&](auto rc, auto cc, T after) -> T {
      constexpr int r = decltype(rc)::value, c = decltype(cc)::value;
      if (MODE == 1 && r >= 2 && remote)
        return ld_cluster(rb + (unsigned)((tile4_off<2, W>(r, c) - W * Sh::kTileStride) * esz), T(0));
      return lds_after(B + tile4_off<OFF, W>(r, c), after);
    };
#define JPD_T4(r, c, after) t[4 * (r) + (c)] = elem(std::integral_constant<int, (r)>{}, std::integral_constant<int, (c)>{}, after)
    if (MODE != 2) {
      // The entries and the (c,s) pairs are loaded group by group, each load pinned behind a result of the group
      // before it (lds_after): the kernel runs at the 128-register cap with 64 registers of eigenvector rows, and
      // with all 16 entries and eight (c,s) doubles loaded up front ptxas spills an eigenvector row in every step
      // -- a local-memory store that the MEMBAR.ALL.GPU of the cluster barrier then waits for (measured: +400
      // cycles per barrier, tests/scratch/t4prof.py).
      const T zero = T(0);
      const T2 p0 = csP[0], q0 = csQ[0];
      if (MODE == 1) {
        // pattern B: the lower half may come from the neighbour CTA -- all eight loads go out at once (one round
        // trip instead of four); only the upper half is loaded group by group
        JPD_T4(2, 0, zero); JPD_T4(2, 1, zero); JPD_T4(2, 2, zero); JPD_T4(2, 3, zero);
        JPD_T4(3, 0, zero); JPD_T4(3, 1, zero); JPD_T4(3, 2, zero); JPD_T4(3, 3, zero);
      }
#define JPD_T4L(r, c, after) do { if (MODE != 1) JPD_T4(r, c, after); } while (0)
      JPD_T4(0, 0, zero); JPD_T4(0, 2, zero); JPD_T4L(2, 0, zero); JPD_T4L(2, 2, zero);
      rotate_block4<T, true>(t[0], t[2], t[8], t[10], p0.x, p0.y, q0.x, q0.y);
      const T2 q1 = lds_after(csQ + 1, t[0]);
      JPD_T4(0, 1, t[0]); JPD_T4(0, 3, t[0]); JPD_T4L(2, 1, t[0]); JPD_T4L(2, 3, t[0]);
      rotate_block4<T, true>(t[1], t[3], t[9], t[11], p0.x, p0.y, q1.x, q1.y);
      const T2 p1 = lds_after(csP + 1, t[1]);
      JPD_T4(1, 0, t[1]); JPD_T4(1, 2, t[1]); JPD_T4L(3, 0, t[1]); JPD_T4L(3, 2, t[1]);
      rotate_block4<T, true>(t[4], t[6], t[12], t[14], p1.x, p1.y, q0.x, q0.y);
      JPD_T4(1, 1, t[4]); JPD_T4(1, 3, t[4]); JPD_T4L(3, 1, t[4]); JPD_T4L(3, 3, t[4]);
      rotate_block4<T, true>(t[5], t[7], t[13], t[15], p1.x, p1.y, q1.x, q1.y);
      const T2 p2 = lds_after(csP + 2, t[5]), q2 = lds_after(csQ + 2, t[5]);
      rotate_block4<T, false>(t[0], t[3], t[12], t[15], p2.x, p2.y, q2.x, q2.y);
      const T2 q3 = lds_after(csQ + 3, t[0]);
      rotate_block4<T, false>(t[1], t[2], t[13], t[14], p2.x, p2.y, q3.x, q3.y);
      const T2 p3 = lds_after(csP + 3, t[1]);
      rotate_block4<T, false>(t[4], t[7], t[8], t[11], p3.x, p3.y, q2.x, q2.y);
      rotate_block4<T, false>(t[5], t[6], t[9], t[10], p3.x, p3.y, q3.x, q3.y);
    } else {
      const T zero = T(0);
#pragma unroll
      for (int k = 0; k < 16; ++k) t[k] = B[k];          // the inner step works on aligned, local tiles
      (void)zero;
      const T2 p0 = csP[0], p1 = csP[1], q0 = csQ[0], q1 = csQ[1];
      rotate_block4<T, false>(t[0], t[1], t[4], t[5], p0.x, p0.y, q0.x, q0.y);
      rotate_block4<T, false>(t[2], t[3], t[6], t[7], p0.x, p0.y, q1.x, q1.y);
      rotate_block4<T, false>(t[8], t[9], t[12], t[13], p1.x, p1.y, q0.x, q0.y);
      rotate_block4<T, false>(t[10], t[11], t[14], t[15], p1.x, p1.y, q1.x, q1.y);
    }
#undef JPD_T4L
#undef JPD_T4
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
  // ---- pattern B: rows {0,1} x column block pair j (CTA 0), and row block pair j x columns {m-2, m-1} (the CTA
  // that owns tile row j; its lower two rows may live in the neighbour).  The 64 top strips go to lanes 0..3 of
  // the 16 warps (four neighbouring strips per warp: no bank conflicts), the 16 right strips, of which two wait
  // for a neighbour's memory, stay together in the last half warp.
  if (MODE == 1) {
    const int r = (tid & 31) < 4 ? (tid >> 5) * 4 + (tid & 31) : (tid >= Sh::kThreads - 16 ? 64 + tid - (Sh::kThreads - 16) : 1 << 20);
    if (r < 64) {
      const int j = r;
      if (cx.rank == 0 && j < npairs) {      // top strip j: tile row 0, tiles (0,j) columns 2,3 and (0,j+1) columns 0,1
        T* S = cx.M + Sh::toff(0, j);
        T* s0 = S + 2; T* s1 = S + 3; T* s2 = S + Sh::kTileStride; T* s3 = S + Sh::kTileStride + 1;
        T a0 = s0[0], a1 = s1[0], a2 = s2[0], a3 = s3[0];
        T b0 = s0[4], b1 = s1[4], b2 = s2[4], b3 = s3[4];
        const T2* g = cx.cs() + Sh::kRec * j;
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
        const unsigned rb = remote ? cluster_map(cx.m_u32 + (unsigned)((Sh::toff(j + 1, Qc) + 2) * esz), (unsigned)Sh::owner_row(j + 1)) : 0u;
        // k = 0,1: rows 2,3 of tile (j, Qc); k = 2,3: rows 0,1 of tile (j+1, Qc)
        T a0 = S[8], a1 = S[12], b0 = S[9], b1 = S[13];
        T a2, a3, b2, b3;
        if (remote) { a2 = ld_cluster(rb, T(0)); a3 = ld_cluster(rb + 4 * esz, T(0)); b2 = ld_cluster(rb + esz, T(0)); b3 = ld_cluster(rb + 5 * esz, T(0)); }
        else { T* S2 = S + W * Sh::kTileStride; a2 = S2[0]; a3 = S2[4]; b2 = S2[1]; b3 = S2[5]; }
        const T2* g = cx.cs() + Sh::kRec * j;
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
  const T2* ab = cx.ab();
  const int J0 = 16 * cx.rank + 4 * ech;    // block pair of this chunk's first four rows
  if (MODE == 1) {
    if ((ech > 0 || cx.rank > 0) && J0 - 1 < npairs) {       // the block pair shared with the chunk below
      const T2* h = ab + Sh::kRec * (J0 - 1);
      const T2 h0 = h[0], h1 = h[1], h2 = h[2], h3 = h[3];
#pragma unroll
      for (int col = 0; col < 2; ++col) {
        const int c = cx.c0 + col * (Sh::kN / 2), cb = Sh::kERows * col;
        T y0, y1;
        // chunk below: slot ech - 1; for ech == 0 it is the top chunk of CTA rank - 1, which PUSHED its rows into
        // this CTA's slot 3 (cl4_back) -- a local load here instead of a round trip to the neighbour
        const int below = (ech + 3) & 3;
        y0 = cx.xhi()[(2 * below + 0) * Sh::kN + c]; y1 = cx.xhi()[(2 * below + 1) * Sh::kN + c];
        T lo0 = e[cb], lo1 = e[cb + 1];
        fast_rotate_group<T, T2, MODE>(y0, y1, lo0, lo1, h0, h1, h2, h3);
        e[cb] = lo0; e[cb + 1] = lo1;
      }
    }
    if ((ech < 3 || cx.rank < 3) && J0 + 3 < npairs) {       // the block pair shared with the chunk above
      const T2* h = ab + Sh::kRec * (J0 + 3);
      const T2 h0 = h[0], h1 = h[1], h2 = h[2], h3 = h[3];
#pragma unroll
      for (int col = 0; col < 2; ++col) {
        const int c = cx.c0 + col * (Sh::kN / 2), cb = Sh::kERows * col;
        T y2, y3;
        const int above = (ech + 1) & 3;                       // for ech == 3: pushed by the bottom chunk of CTA rank + 1
        y2 = cx.xlo()[(2 * above + 0) * Sh::kN + c]; y3 = cx.xlo()[(2 * above + 1) * Sh::kN + c];
        T hi0 = e[cb + 14], hi1 = e[cb + 15];
        fast_rotate_group<T, T2, MODE>(hi0, hi1, y2, y3, h0, h1, h2, h3);
        e[cb + 14] = hi0; e[cb + 15] = hi1;
      }
    }
  }
#pragma unroll
  for (int jj = 0; jj < 4 - (MODE == 1 ? 1 : 0); ++jj) {
    if (J0 + jj < npairs) {
      const T2* h = ab + Sh::kRec * (J0 + jj);
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

// first phase of a block step: the pivots this CTA owns (16 threads of warp 0).  The eigenvector rotations of a
// step run in its SECOND phase, beside the tiles (cl4_back): measured with clock64 (tests/scratch/t4prof.py), the
// pivot chain takes 1.4 times longer (1990 against 1410 cycles) when the other warps of its scheduler do eigenvector
// work at the same time, and
// the pivot warp's own eigenvector work then follows it on the critical path.
template <typename T, bool EVECS, int MODE>
__device__ __forceinline__ void cl4_front(const Cl4Ctx<T>& cx) {
  using Sh = Cl4Shape;
  const int npairs = (MODE == 1) ? cx.nbp - 1 : cx.nbp;
  if (threadIdx.x < 16) {
    const int J = Sh::global_row(threadIdx.x, cx.rank);
    if (J < npairs) cl4_diag<T, EVECS, MODE>(cx, J);
  }
}

// Pivots published: warp 0 sends this CTA's two ranges of pivot records -- block pairs 8 rank .. 8 rank + 7 and
// 8 (7 - rank) .. + 7, the tile rows it owns -- to the three other CTAs (6 bulk copies; their cost is per copy,
// about 90 cycles each, which is why (c,s) and (alpha,beta) share a record); every thread then waits until the
// ranges of the three others have landed here and the CTA's own pivot warp has arrived.  Records of block pairs
// that do not exist at this n travel too (never read).
template <typename T>
__device__ __forceinline__ void cl4_exchange_pivots(const Cl4Ctx<T>& cx, unsigned& phase) {
  using Sh = Cl4Shape;
  constexpr unsigned e2 = 2 * (unsigned)sizeof(T);
  constexpr unsigned kRange = 8 * Sh::kRec * e2;
  if (threadIdx.x < 32) {
    fence_proxy_async_smem();     // the pivot lanes' plain stores, before the bulk copies read them
    __syncwarp();
    const int lane = threadIdx.x;
    if (lane < 6) {
      const unsigned dst = (unsigned)((cx.rank + 1 + (lane >> 1)) & 3);
      const int J0 = (lane & 1) ? 8 * (7 - cx.rank) : 8 * cx.rank;
      const unsigned src = cx.cs_u32() + (unsigned)(J0 * Sh::kRec) * e2;
      bulk_copy_to_peer(cluster_map(src, dst), src, kRange, cluster_map(cx.bar_u32(), dst));
    }
    if (lane == 0) mbar_arrive_expect_tx(cx.bar_u32(), 3 * 2 * kRange);   // release: warp 0's stores, ordered by __syncwarp
  }
  mbar_wait(cx.bar_u32(), phase);
  phase ^= 1u;
}

// second phase of a block step: the tiles this CTA owns and the step's eigenvector rotations.  The tile work is
// bound by shared-memory wavefronts, the eigenvector rows by the FP64 pipe: odd warps take them in the other order
// so that the two kinds of work overlap on every scheduler.  After a pattern-A step the rows next to the chunk
// boundaries are left in the exchange buffers for the pattern-B step that follows (the cluster barrier behind
// this phase orders them).
template <typename T, bool EVECS, int MODE>
__device__ __forceinline__ void cl4_back(const Cl4Ctx<T>& cx, T (&e)[EVECS ? 32 : 1]) {
  using Sh = Cl4Shape;
  if constexpr (!EVECS) {
    cl4_tiles<T, MODE>(cx);
  } else {
    const bool evecs_first = (threadIdx.x >> 5) & 1;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {        // one copy of each body (the loop is not unrolled: instruction cache)
      if ((pass == 0) != evecs_first) {
        cl4_tiles<T, MODE>(cx);
      } else {
        cl4_evecs<T, MODE>(cx, e);
        if constexpr (MODE == 0) {
#pragma unroll
          for (int col = 0; col < 2; ++col) {
            const int c = cx.c0 + col * (Sh::kN / 2), cb = Sh::kERows * col;
            // slot ech of this CTA -- except the rows a NEIGHBOUR CTA needs: the bottom chunk pushes its first two
            // rows into slot 0 of CTA rank - 1, the top chunk its last two into slot 3 of CTA rank + 1 (four remote
            // stores here instead of four remote loads, with their round trip, in the pattern-B step)
            constexpr int esz = (int)sizeof(T);
            const unsigned lo = (unsigned)(((2 * cx.ech + 0) * Sh::kN + c) * esz), hi = lo;
            if (cx.ech == 0) {
              if (cx.rank > 0) {
                const unsigned a = cluster_map(cx.xlo_u32() + lo, (unsigned)(cx.rank - 1));
                st_cluster(a, e[cb]); st_cluster(a + (unsigned)(Sh::kN * esz), e[cb + 1]);
              }
            } else {
              cx.xlo()[(2 * cx.ech + 0) * Sh::kN + c] = e[cb]; cx.xlo()[(2 * cx.ech + 1) * Sh::kN + c] = e[cb + 1];
            }
            if (cx.ech == 3) {
              if (cx.rank < 3) {
                const unsigned a = cluster_map(cx.xhi_u32() + hi, (unsigned)(cx.rank + 1));
                st_cluster(a, e[cb + 14]); st_cluster(a + (unsigned)(Sh::kN * esz), e[cb + 15]);
              }
            } else {
              cx.xhi()[(2 * cx.ech + 0) * Sh::kN + c] = e[cb + 14]; cx.xhi()[(2 * cx.ech + 1) * Sh::kN + c] = e[cb + 15];
            }
          }
        }
      }
    }
  }
}

// One block step.  The barrier behind the second phase is a plain barrier.cluster arrive.release / wait.acquire:
// tiles that straddle two CTAs, the pivots' entries in a neighbour's diagonal tile and the boundary rows of E are
// read by other CTAs in the next step.  (The relaxed arrive with the release spelled out as fence.acq_rel.cluster in
// the threads that wrote remotely was measured too and is not faster: a sixteenth of the threads of EVERY warp
// write into a neighbour in the tile phase, so every warp pays its fence anyway.)
template <typename T, bool EVECS, int MODE>
__device__ __forceinline__ void cl4_step(const Cl4Ctx<T>& cx, T (&e)[EVECS ? 32 : 1], unsigned& phase) {
  cl4_front<T, EVECS, MODE>(cx);
  cl4_exchange_pivots<T>(cx, phase);
  cl4_back<T, EVECS, MODE>(cx, e);
  cluster_barrier();
}

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
  T2* ab = cs + 4;                 // same records, second half
  T2* fr = cs + Sh::kRec * Sh::kBP;
  T* dg = reinterpret_cast<T*>(fr + Sh::kN);
  T* xlo = dg + Sh::kN;
  T* xhi = xlo + 2 * Sh::kEChunks * Sh::kN;
  int* perm = reinterpret_cast<int*>(smem_raw + Sh::table_bytes() + ((size_t)Sh::kElems * sizeof(T) + 15) / 16 * 16);
  int* flags = perm + Sh::kN;      // [4]
  int* rots = flags + 4;           // [4]
  int* ovf = rots + 4;             // scratch of the table builder: [2]  (tables are final before use)
  unsigned phase = 0;              // parity of the mbarrier phase the next block step waits for

  Cl4Ctx<T> cx;
  cx.M = M; cx.m_u32 = smem_u32(M);
  cx.m = m; cx.nbp = nbp; cx.rank = rank; cx.c0 = tid & 127; cx.ech = tid >> 7;
  if (tid == 0) { mbar_init(cx.bar_u32(), 1); fence_mbar_init_cluster(); }   // the cluster barrier below publishes it

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
    // The tile rows of pattern B that reach into a neighbour CTA go first, so that their tiles land in the first
    // slots of every class, i.e. in the first two warps: a straddling tile costs a round trip to the neighbour and
    // eight remote stores (about 1000 cycles per warp that holds one), which only those warps now pay.
    for (int ll = 0; ll < 2 * Sh::kRowsLocal; ++ll) {
      const int l = ll % Sh::kRowsLocal;
      const int P = Sh::global_row(l, rank);
      if (P >= np) continue;
      const unsigned remote = (pat == 1 && Sh::owner_row(P + 1) != rank) ? 1u : 0u;
      if ((ll < Sh::kRowsLocal) != (remote != 0)) continue;
      for (int Q = P + 1 + ((cl - l - P - 1) & 15); Q < np; Q += 16) {     // (l + Q) mod 16 == cl
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
    int sweeps = 0;
    if (tid == 0) *cx.rot_count() = 0;     // ordered before the first pivots by the cluster barriers below
    bool need = false;
    for (;;) {
      // every CTA gets a copy of the diagonal ...
      if (tid < 4 * Sh::kRowsLocal) {
        const int l = tid >> 2, k = tid & 3;
        const int i = 4 * Sh::global_row(l, rank) + k;
        if (i < m) {
          const T d = M[(l * Sh::kW + Sh::global_row(l, rank)) * Sh::kTileStride + 5 * k];
#pragma unroll
          for (unsigned r = 0; r < 4; ++r) st_cluster(cluster_map(cx.dg_u32() + (unsigned)(i * esz), r), d);
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

      cl4_step<T, EVECS, 2>(cx, e, phase);
      for (int bs = 0; bs < nb; bs += 2) {
        cl4_step<T, EVECS, 0>(cx, e, phase);
        cl4_step<T, EVECS, 1>(cx, e, phase);
      }
      ++sweeps;
      if constexpr (EVECS) {
#pragma unroll
        for (int k = 0; k < Sh::kERows; ++k) {   // fold the row scales back into the rows
          const int pk = 64 * rank + Sh::kERows * cx.ech + k;     // its scale lives with tile row pk / 4
          const T fk = ld_cluster(cluster_map(cx.fr_u32() + (unsigned)(pk * 2 * esz), (unsigned)Sh::owner_row(pk >> 2)), T(0));
          e[k] *= fk; e[Sh::kERows + k] *= fk;
        }
        cluster_barrier();   // every CTA is done with its neighbours' boundary rows and with the scales
        if (tid < Sh::kN) { T2 one; one.x = T(1); one.y = T(1); fr[tid] = one; }
      }
    }
    const bool converged = !need && max_num_sweeps > 0;
    // rotations: block sum, then every CTA tells CTA 0
    if (tid == 0) st_cluster_i32(cluster_map(smem_u32(rots) + (unsigned)(rank * 4), 0u), *cx.rot_count());

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
