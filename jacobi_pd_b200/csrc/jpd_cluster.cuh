// jpd_cluster.cuh -- tier 4: one 4-CTA thread-block cluster per matrix, 129 <= n <= 256.
//
// Path being replaced: Jacobi::Diagonalize, /root/reference/include/jacobi_pd.hpp:159-220
// (CalcRot :233-256, ApplyRot :327-393, ApplyRotLeft :409-419, SortRows :477-508).
//
// The tier-3 scheme (jpd_cta.cuh) no longer fits one SM at n = 256: E is 512 KB (an SM's
// register file holds 256 KB) and the live upper triangle of M 265 KB (227 KB of shared
// memory).  A cluster of 4 CTAs x 512 threads has both:
//  * E in REGISTERS: CTA c owns positions 64c .. 64c+63, thread (v, chl) rows
//    64c + 32 chl .. + 31 of column v.  Rows that pair across a chunk boundary in pattern B are
//    pushed straight into the consumer CTA's receive buffer (st.shared::cluster).
//  * M in DISTRIBUTED SHARED MEMORY: rows are dealt to the CTAs in groups of 16 in snake
//    order (group g of a 64-row super group goes to CTA g or 3-g), which gives every CTA the
//    same number of 2x2 blocks of the upper triangle (2032 of 8128 at n = 256) and 64 rows
//    = 133 KB of storage.  A CTA updates the blocks whose first row it owns; in pattern A both
//    rows are local, in pattern B one pair in eight has its second row in a neighbour CTA and
//    is reached through ld/st.shared::cluster.  Same even/odd split columns, row pitch
//    == 3 (mod 8) and bank-class schedule tables as tiers 2/3 (conflict-free local accesses).
//  * Every CTA holds all (c,s) pairs: four threads per pair compute the rotation and thread k
//    of the quad writes it into CTA k (triple-buffered by step).
//  * Barriers: measured on B200, barrier.cluster arrive.release/wait.acquire costs 443 cycles
//    (a MEMBAR.ALL.GPU per warp), the relaxed form 106, __syncthreads 45.  The inner loop
//    therefore uses relaxed arrives and spells the release out as fence.acq_rel.cluster in the
//    threads that wrote another CTA's memory plus, where a neighbour will read this CTA's own
//    rows, ONE cumulative fence per CTA after the block barrier (cluster_arrive_light); 3 cluster
//    barriers per A/B step pair (the barrier after a pattern-A block phase is block-local); a
//    step's eigenvector phase runs during the NEXT step's pivot phase (between arrive and wait
//    of its first barrier).
// Same odd-even ordering, skip test, rotation, convergence rule and return value.
#pragma once
#include "jpd_common.cuh"
#include "jpd_warp.cuh"

namespace jpd {

struct ClusterShape {
  static constexpr int kN = 256;                 // largest matrix
  static constexpr int kCtas = 4;
  static constexpr int kThreads = 512;
  static constexpr int kRowsLocal = 64;
  static constexpr int kH = kN / 2;
  static constexpr int kLd = kN + 3;             // 259 == 3 (mod 8)
  static constexpr int kMElems = kRowsLocal * kLd;
  static constexpr int kPairsMax = kN / 2;
  static constexpr int kIt = 4;                  // block iterations per step (tests/test_schedule.py)
  static constexpr int kTblWords = 2 * kIt * kThreads;   // two words per entry, per pattern
  // elements of T: M | c[3][128] | s[3][128] | eigenvalues | rows from the chunk below / above
  static constexpr int kCsBufs = 3;              // (c,s) arrays are triple-buffered by step
  static constexpr int kElems = kMElems + kCsBufs * 2 * kPairsMax + kN + 4 * kN;
  static constexpr int kInts = kN + 8;           // rank[] | flags[4] | rot_total | pad
  JPD_CX static constexpr size_t table_bytes() { return (size_t)2 * kTblWords * sizeof(unsigned); }
  JPD_CX static constexpr size_t block_bytes(size_t elem) {
    return table_bytes() + ((size_t)kElems * elem + 15) / 16 * 16 + (size_t)kInts * sizeof(int);
  }
  JPD_HD static int owner(int r) { const int g = (r >> 4) & 3; return ((r >> 6) & 1) ? 3 - g : g; }
  JPD_HD static int lrow(int r) { return ((r >> 6) << 4) | (r & 15); }
  JPD_HD static int slot(int c) { return (c >> 1) + (c & 1) * kH; }
  // global row of local row lr in CTA `rank`
  JPD_HD static int grow(int lr, int rank) {
    const int sg = lr >> 4;
    return 64 * sg + 16 * ((sg & 1) ? 3 - rank : rank) + (lr & 15);
  }
};

#if defined(__CUDACC__)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned cluster_map(unsigned addr, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ unsigned cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_barrier() { cluster_arrive(); cluster_wait(); }
// Light cluster barrier for the inner loop.  barrier.cluster.arrive.release costs every warp a
// MEMBAR.ALL.GPU (r1 profile: membar + barrier = 3 stalled warps per issue), so the release is
// spelled out as fences in the threads that need one:
//  * a thread that wrote ANOTHER CTA's shared memory since the last barrier fences its own
//    stores (wrote_remote);
//  * if another CTA will READ, after this barrier, shared memory of THIS CTA that was written
//    with plain local stores (publish_local: rows of M that pattern B reaches through
//    ld.shared::cluster, the diagonal read by a neighbour's pivot, the convergence test),
//    one thread fences AFTER the block barrier: fence.acq_rel.cluster is cumulative, so it
//    covers every store of the CTA that the block barrier ordered before it, and that
//    thread's (relaxed) arrive is then the release the readers' wait.acquire pairs with.
// barrier.cluster.wait is an acquire by default (PTX ISA: .acquire is assumed when absent).
__device__ __forceinline__ void cluster_fence() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
__device__ __forceinline__ void cluster_arrive_light(bool wrote_remote, bool publish_local) {
  if (wrote_remote) cluster_fence();
  __syncthreads();
  if (publish_local && threadIdx.x == 0) cluster_fence();
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_light() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ double ld_cluster(unsigned a, double) {
  double v; asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory"); return v;
}
__device__ __forceinline__ float ld_cluster(unsigned a, float) {
  float v; asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory"); return v;
}
__device__ __forceinline__ void st_cluster(unsigned a, double v) {
  asm volatile("st.shared::cluster.f64 [%0], %1;" :: "r"(a), "d"(v) : "memory");
}
__device__ __forceinline__ void st_cluster(unsigned a, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" :: "r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ int ld_cluster_i32(unsigned a) {
  int v; asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v;
}
__device__ __forceinline__ void st_cluster_i32(unsigned a, int v) {
  asm volatile("st.shared::cluster.s32 [%0], %1;" :: "r"(a), "r"(v) : "memory");
}

// cluster-space byte address of element (r,c); m_u32 = this CTA's shared address of M
template <typename T>
__device__ __forceinline__ unsigned cl_eaddr(unsigned m_u32, int r, int c) {
  using Sh = ClusterShape;
  return cluster_map(m_u32 + (unsigned)((Sh::lrow(r) * Sh::kLd + Sh::slot(c)) * (int)sizeof(T)), (unsigned)Sh::owner(r));
}

// pair handled by pivot lane t (0..31) of CTA `rank`: the pairs whose first row the CTA owns
__device__ __forceinline__ int cl_pair_of(int t, int rank) {
  const int sg = t >> 3;                                   // 64-row super group
  const int grp = 4 * sg + ((sg & 1) ? 3 - rank : rank);   // 16-row group = 8 pairs
  return 8 * grp + (t & 7);
}

// entry [it*512 + tid]: word 0 = local offset of the block's first element | K << 16 | L << 23 |
// (second row lives in another CTA) << 30 (0xffffffff = idle); word 1 = first element of the second
// row: local element offset, or cluster-space byte address when remote
template <typename T>
__device__ __forceinline__ void cluster_build_schedule(unsigned* tbl, unsigned m_u32, int npr, int off, int rank) {
  using Sh = ClusterShape;
  for (int i = threadIdx.x; i < Sh::kTblWords; i += Sh::kThreads) tbl[i] = 0xffffffffu;
  __syncthreads();
  if (threadIdx.x < 16) {
    const int cl = threadIdx.x;
    int slot = 0;
    for (int K = 0; K < npr; ++K) {
      const int p = 2 * K + off;
      if (Sh::owner(p) != rank) continue;
      for (int L = (cl - 6 * (K & 7)) & 15; L < npr; L += 16) {
        if (L > K) {
          const int idx = (slot >> 5) * Sh::kThreads + (slot & 31) * 16 + cl;
          if (idx < Sh::kIt * Sh::kThreads) {
            const bool remote = Sh::owner(p + 1) != rank;
            tbl[2 * idx] = (unsigned)(Sh::lrow(p) * Sh::kLd + Sh::slot(2 * L + off)) | ((unsigned)K << 16) |
                           ((unsigned)L << 23) | (remote ? 1u << 30 : 0u);
            tbl[2 * idx + 1] = remote ? cl_eaddr<T>(m_u32, p + 1, 2 * L + off)
                                      : (unsigned)(Sh::lrow(p + 1) * Sh::kLd + Sh::slot(2 * L + off));
          }
          ++slot;
        }
      }
    }
  }
}

struct ClusterPtrs {      // this CTA's shared-window addresses of the arrays other CTAs write
  unsigned m, cc, ss, below, above, flags;
};

// Everything a step needs, bundled (all threads of the CTA hold the same values except v, chl)
template <typename T> struct ClusterCtx {
  T* M; T* cc; T* ss; T* from_below; T* from_above;
  ClusterPtrs sp;
  const unsigned* tblA; const unsigned* tblB;
  int m, v, chl, rank;
};

// ---- phase 1: four threads per pair whose first row this CTA owns (all four compute the
// rotation; lane k of the quad publishes (c,s) to CTA k, lane 0 updates the pair's entries).
// Returns whether this thread wrote another CTA's shared memory.
template <typename T, int OFF>
__device__ __forceinline__ bool cluster_pivot(const ClusterCtx<T>& cx, int buf, int& n_rot) {
  using Sh = ClusterShape;
  constexpr int o01 = (OFF == 0) ? Sh::kH : 1 - Sh::kH;
  constexpr int esz = (int)sizeof(T);
  const int tid = threadIdx.x;
  const int npr = (cx.m >> 1) - OFF;
  // (c,s) arrays are triple-buffered by step: the eigenvector phase of step t runs during
  // step t+1, and a CTA that is ahead may already publish step t+2 (only the cluster
  // barrier after a pattern-B step holds it back); it cannot reach step t+3 before every
  // CTA has arrived at the first barrier of step t+2, i.e. finished step t+1 completely
  const unsigned cc_u = cx.sp.cc + (unsigned)(buf * Sh::kPairsMax * esz);
  const unsigned ss_u = cx.sp.ss + (unsigned)(buf * Sh::kPairsMax * esz);
  T* M = cx.M;
  bool wrote_remote = false;
  if (tid < 128) {
    const int j = cl_pair_of(tid >> 2, cx.rank);
    const unsigned target = (unsigned)(tid & 3);
    const bool act = j < npr;
    const int p = 2 * j + OFF;
    T* pp = M + Sh::lrow(p) * Sh::kLd + Sh::slot(p);        // (p,p) and (p,p+1): this CTA's row
    T* pq = pp + o01;
    T* qq = M + Sh::lrow(p + 1) * Sh::kLd + Sh::slot(p + 1);
    const bool q_remote = (OFF == 1) && Sh::owner(p + 1) != cx.rank;
    unsigned aqq = 0;
    T mpp = T(0), a = T(0), mqq = T(0);
    if (act) {
      aqq = cl_eaddr<T>(cx.sp.m, p + 1, p + 1);
      mpp = *pp; a = *pq;
      mqq = q_remote ? ld_cluster(aqq, T(0)) : *qq;
    }
    __syncwarp();                         // every quad has read its old entries
    if (act) {
      T c, s, npp, nqq;
      const bool rot = pivot_exchange_rotation(mpp, mqq, a, c, s, npp, nqq);
      st_cluster(cluster_map(cc_u + (unsigned)(j * esz), target), c);
      st_cluster(cluster_map(ss_u + (unsigned)(j * esz), target), s);
      wrote_remote = true;
      if (target == 0) {
        if (rot) { *pq = T(0); ++n_rot; }
        *pp = npp;
        if (q_remote) st_cluster(aqq, nqq); else *qq = nqq;
      }
    }
  }
  return wrote_remote;
}

// ---- phase 2a: the 2x2 blocks whose first row this CTA owns (+ the end strips of pattern B)
template <typename T, int OFF>
__device__ __forceinline__ bool cluster_blocks(const ClusterCtx<T>& cx, int buf) {
  using Sh = ClusterShape;
  constexpr int o01 = (OFF == 0) ? Sh::kH : 1 - Sh::kH;
  constexpr int esz = (int)sizeof(T);
  const int tid = threadIdx.x;
  const int npr = (cx.m >> 1) - OFF;
  T* M = cx.M;
  const T* cc = cx.cc + buf * Sh::kPairsMax;
  const T* ss = cx.ss + buf * Sh::kPairsMax;
  const unsigned* tbl = OFF == 0 ? cx.tblA : cx.tblB;
  bool wrote_remote = false;
#pragma unroll
  for (int it = 0; it < Sh::kIt; ++it) {
    const uint2 ent = reinterpret_cast<const uint2*>(tbl)[it * Sh::kThreads + tid];
    if (ent.x != 0xffffffffu) {
      T* b = M + (ent.x & 0xffffu);
      const int K = (ent.x >> 16) & 0x7fu, L = (ent.x >> 23) & 0x7fu;
      const bool remote = (OFF == 1) && ((ent.x >> 30) & 1u);   // pattern A pairs never leave their CTA
      const T ck = cc[K], sk = ss[K], cl = cc[L], sl = ss[L];
      T b00 = b[0], b01 = b[o01];
      T b10, b11;
      T* b1 = M + ent.y;
      const unsigned a10 = ent.y, a11 = ent.y + (unsigned)(o01 * esz);
      if (remote) { b10 = ld_cluster(a10, T(0)); b11 = ld_cluster(a11, T(0)); }
      else { b10 = b1[0]; b11 = b1[o01]; }
      exchange_rotate_block(b00, b01, b10, b11, ck, sk, cl, sl);
      b[0] = b00;
      b[o01] = b01;
      if (remote) { st_cluster(a10, b10); st_cluster(a11, b11); wrote_remote = true; }
      else { b1[0] = b10; b1[o01] = b11; }
    }
  }
  // pattern B: row 0 x column pairs (CTA owning row 0), row pairs x column m-1
  if (OFF == 1) {
    if (cx.rank == Sh::owner(0) && tid >= 64 && tid < 64 + Sh::kPairsMax) {
      const int j = tid - 64;
      if (j < npr) {
        T* x0 = M + Sh::lrow(0) * Sh::kLd + Sh::slot(2 * j + 1);
        T v0 = x0[0], v1 = x0[o01];
        exchange_rotate_pair(v0, v1, cc[j], ss[j]);
        x0[0] = v0; x0[o01] = v1;
      }
    } else if (tid >= 32 && tid < 64) {
      const int j = cl_pair_of(tid - 32, cx.rank);
      if (j < npr) {
        const unsigned a0 = cl_eaddr<T>(cx.sp.m, 2 * j + 1, cx.m - 1), a1 = cl_eaddr<T>(cx.sp.m, 2 * j + 2, cx.m - 1);
        T v0 = ld_cluster(a0, T(0)), v1 = ld_cluster(a1, T(0));
        exchange_rotate_pair(v0, v1, cc[j], ss[j]);
        st_cluster(a0, v0);
        st_cluster(a1, v1);
        wrote_remote = true;
      }
    }
  }
  return wrote_remote;
}

// ---- phase 2b: eigenvector rows of a step of pattern OFF (registers; reads only (c,s) and,
// for pattern B, the two rows received from the neighbouring chunks)
// recv_above / recv_below: the rows received from the neighbouring chunks (pattern B only),
// read by the caller BEFORE it arrives at the cluster barrier that lets the neighbours run
// ahead and push the next pair of rows (cluster_recv_rows).
template <typename T, int OFF>
__device__ __forceinline__ void cluster_evecs(const ClusterCtx<T>& cx, int buf, T (&e)[32],
                                              T recv_above = T(0), T recv_below = T(0)) {
  using Sh = ClusterShape;
  const int npr = (cx.m >> 1) - OFF;
  const int g = 2 * cx.rank + cx.chl;
  const T* cc = cx.cc + buf * Sh::kPairsMax;
  const T* ss = cx.ss + buf * Sh::kPairsMax;
  const int j0 = 16 * g;
  T new_lo = e[0], new_hi = e[31];
  if (OFF == 1) {
    if (g > 0 && j0 - 1 < npr) {
      const T rc = cc[j0 - 1], rs = ss[j0 - 1];
      new_lo = jfma(rc, recv_above, -(rs * e[0]));
    }
    if (g < 7 && j0 + 15 < npr) {
      const T rc = cc[j0 + 15], rs = ss[j0 + 15];
      new_hi = jfma(rs, e[31], rc * recv_below);
    }
  }
#pragma unroll
  for (int jj = 0; jj < 16 - OFF; ++jj) {
    const int p = 2 * jj + OFF, q = p + 1;
    if (j0 + jj < npr) {
      exchange_rotate_pair(e[p], e[q], cc[j0 + jj], ss[j0 + jj]);
    }
  }
  if (OFF == 1) { e[0] = new_lo; e[31] = new_hi; }
}

// hand the rows that pair across a chunk boundary in pattern B to their consumers: chunk g-1
// needs row 32g, chunk g+1 needs row 32g+31.  Returns whether the push left this CTA.
template <typename T>
__device__ __forceinline__ bool cluster_push_rows(const ClusterCtx<T>& cx, const T (&e)[32]) {
  using Sh = ClusterShape;
  constexpr int esz = (int)sizeof(T);
  const int g = 2 * cx.rank + cx.chl;
  bool remote = false;
  if (g > 0) {
    const int idx = ((g - 1) & 1) * Sh::kN + cx.v;
    if (cx.chl == 0) { st_cluster(cluster_map(cx.sp.below + (unsigned)(idx * esz), (unsigned)((g - 1) >> 1)), e[0]); remote = true; }
    else cx.from_below[idx] = e[0];
  }
  if (g < 7) {
    const int idx = ((g + 1) & 1) * Sh::kN + cx.v;
    if (cx.chl == 1) { st_cluster(cluster_map(cx.sp.above + (unsigned)(idx * esz), (unsigned)((g + 1) >> 1)), e[31]); remote = true; }
    else cx.from_above[idx] = e[31];
  }
  return remote;
}

template <typename T>
__device__ __forceinline__ void cluster_recv_rows(const ClusterCtx<T>& cx, T& recv_above, T& recv_below) {
  recv_above = cx.from_above[cx.chl * ClusterShape::kN + cx.v];
  recv_below = cx.from_below[cx.chl * ClusterShape::kN + cx.v];
}

// One A step + one B step.  The eigenvector phase of a step is deferred to the NEXT step and
// runs between arrive and wait of that step's first cluster barrier, i.e. while the pivot
// warps compute and publish the next rotations (the longest serial stretch of a step).
//   A: pivot_A | arrive | E(previous B) | wait | blocks_A | block barrier (all writes local)
//   B: pivot_B | arrive | E(this A), push boundary rows | wait | blocks_B | cluster barrier
template <typename T, bool EVECS>
__device__ __forceinline__ void cluster_step_pair(const ClusterCtx<T>& cx, T (&e)[EVECS ? 32 : 1],
                                                  bool have_prev, bool last_of_sweep, int& buf, int& n_rot) {
  const int buf_prev = buf == 0 ? 2 : buf - 1;          // previous B step
  const int buf_a = buf, buf_b = buf == 2 ? 0 : buf + 1;
  bool wr = cluster_pivot<T, 0>(cx, buf_a, n_rot);
  // The rows pushed during the previous B step are taken out of the receive buffers BEFORE the
  // arrive: a neighbour that runs ahead overwrites them (cluster_push_rows of THIS pair) only
  // after it has passed this barrier, i.e. after every CTA has read them.
  T recv_above = T(0), recv_below = T(0);
  if constexpr (EVECS) { if (have_prev) cluster_recv_rows<T>(cx, recv_above, recv_below); }
  cluster_arrive_light(wr, true);    // pivot_B of a neighbour reads this CTA's diagonal
  if constexpr (EVECS) { if (have_prev) cluster_evecs<T, 1>(cx, buf_prev, e, recv_above, recv_below); }
  cluster_wait_light();
  cluster_blocks<T, 0>(cx, buf_a);
  __syncthreads();

  wr = cluster_pivot<T, 1>(cx, buf_b, n_rot);
  cluster_arrive_light(wr, true);    // blocks_B of a neighbour reads rows this CTA wrote in blocks_A
  wr = false;
  if constexpr (EVECS) { cluster_evecs<T, 0>(cx, buf_a, e); wr = cluster_push_rows<T>(cx, e); }
  cluster_wait_light();
  wr = cluster_blocks<T, 1>(cx, buf_b) || wr;
  cluster_arrive_light(wr, last_of_sweep);   // the convergence test reads every CTA's rows
  cluster_wait_light();
  buf = buf_b == 2 ? 0 : buf_b + 1;
}

template <typename T, bool EVECS>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(512, 1)
jpd_cluster_kernel(const T* __restrict__ mat, T* __restrict__ evals, T* __restrict__ evecs,
                   int* __restrict__ n_iters, int n, long long batch, int soa, int sort_criteria,
                   int max_num_sweeps) {
  using Sh = ClusterShape;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int v = tid & 255, chl = tid >> 8;
  const int rank = (int)cluster_ctarank();
  const int g = 2 * rank + chl;
  const int m = n + (n & 1);
  constexpr int esz = (int)sizeof(T);

  unsigned* tblA = reinterpret_cast<unsigned*>(smem_raw);
  unsigned* tblB = tblA + Sh::kTblWords;
  T* M = reinterpret_cast<T*>(smem_raw + Sh::table_bytes());
  T* cc = M + Sh::kMElems;
  T* ss = cc + Sh::kCsBufs * Sh::kPairsMax;
  T* ev = ss + Sh::kCsBufs * Sh::kPairsMax;
  T* from_below = ev + Sh::kN;
  T* from_above = from_below + 2 * Sh::kN;
  int* rank_of = reinterpret_cast<int*>(smem_raw + Sh::table_bytes() + ((size_t)Sh::kElems * sizeof(T) + 15) / 16 * 16);
  int* flags = rank_of + Sh::kN;
  int* rot_total = flags + 4;
  ClusterPtrs sp;
  sp.m = smem_u32(M); sp.cc = smem_u32(cc); sp.ss = smem_u32(ss);
  sp.below = smem_u32(from_below); sp.above = smem_u32(from_above); sp.flags = smem_u32(flags);

  ClusterCtx<T> cx;
  cx.M = M; cx.cc = cc; cx.ss = ss; cx.from_below = from_below; cx.from_above = from_above; cx.sp = sp;
  cx.tblA = tblA; cx.tblB = tblB; cx.m = m; cx.v = v; cx.chl = chl; cx.rank = rank;

  cluster_build_schedule<T>(tblA, sp.m, m >> 1, 0, rank);
  cluster_build_schedule<T>(tblB, sp.m, (m >> 1) - 1, 1, rank);
  cluster_barrier();

  const long long n_clusters = gridDim.x / Sh::kCtas;
  for (long long b = blockIdx.x / Sh::kCtas; b < batch; b += n_clusters) {
    // ---- load the rows this CTA owns: upper triangle only (jacobi_pd.hpp:167-171)
    for (int i = tid; i < Sh::kMElems; i += Sh::kThreads) M[i] = T(0);
    if (tid == 0) *rot_total = 0;
    __syncthreads();
    for (int idx = tid; idx < Sh::kRowsLocal * n; idx += Sh::kThreads) {
      const int lr = idx / n, c = idx - lr * n;
      const int r = Sh::grow(lr, rank);
      if (r < n && c >= r) {
        const size_t e_idx = (size_t)r * n + c;
        const T val = soa ? mat[e_idx * (size_t)batch + (size_t)b] : mat[(size_t)b * n * n + e_idx];
        M[lr * Sh::kLd + Sh::slot(c)] = val;
      }
    }
    T e[EVECS ? 32 : 1];
    if (EVECS) {
#pragma unroll
      for (int k = 0; k < 32; ++k) e[k] = (32 * g + k == v) ? T(1) : T(0);   // :172-178
    }
    cluster_barrier();

    // ---- sweeps
    int n_rot = 0, sweeps = 0, buf = 0;
    bool need = false;
    for (;;) {
      // the reference's convergence test (:191) on every pair (i, i+d mod m), 8 threads per i
      int mine = 0;
      if (v < m) {
        const T dii = ld_cluster(cl_eaddr<T>(sp.m, v, v), T(0));
        for (int d = 1 + g; d <= (m >> 1); d += 8) {
          int j = v + d; if (j >= m) j -= m;
          const int lo = v < j ? v : j, hi = v < j ? j : v;
          mine |= pair_needs_rotation(ld_cluster(cl_eaddr<T>(sp.m, lo, hi), T(0)), dii,
                                      ld_cluster(cl_eaddr<T>(sp.m, j, j), T(0))) ? 1 : 0;
        }
      }
      const int cta_need = __syncthreads_or(mine);
      if (tid < 4) st_cluster_i32(cluster_map(sp.flags + (unsigned)(rank * 4), (unsigned)tid), cta_need);
      cluster_arrive_light(tid < 4, false);
      cluster_wait_light();
      need = (flags[0] | flags[1] | flags[2] | flags[3]) != 0;
      if (!need || sweeps >= max_num_sweeps) break;
      for (int step = 0; step < m; step += 2) cluster_step_pair<T, EVECS>(cx, e, step > 0, step + 2 >= m, buf, n_rot);
      if constexpr (EVECS) {   // the last B step's eigenvector phase (no push can follow before the next vote)
        T recv_above, recv_below;
        cluster_recv_rows<T>(cx, recv_above, recv_below);
        cluster_evecs<T, 1>(cx, buf == 0 ? 2 : buf - 1, e, recv_above, recv_below);
      }
      ++sweeps;
    }
    const bool converged = !need && max_num_sweeps > 0;
    if (n_rot) atomicAdd(rot_total, n_rot);

    // ---- eigenvalues (:207-209): every CTA gathers all n, ranks them (same key order as SortRows)
    const int first = ((n & 1) && (sweeps & 1)) ? 1 : 0;
    T my_ev = T(0);
    if (tid < n) { my_ev = ld_cluster(cl_eaddr<T>(sp.m, first + tid, first + tid), T(0)); ev[tid] = my_ev; }
    cluster_barrier();    // also: every CTA's rot_total is final
    if (tid < n) {
      int rk = tid;
      if (sort_criteria >= SORT_DECREASING_EVALS && sort_criteria <= SORT_INCREASING_ABS_EVALS) {
        const bool use_abs = sort_criteria >= SORT_DECREASING_ABS_EVALS;
        const bool decreasing = (sort_criteria == SORT_DECREASING_EVALS) || (sort_criteria == SORT_DECREASING_ABS_EVALS);
        const long long key = sortable_bits(use_abs ? jabs(my_ev) : my_ev);
        rk = 0;
        for (int j = 0; j < n; ++j) {
          const T o = ev[j];
          const long long other = sortable_bits(use_abs ? jabs(o) : o);
          const bool before = decreasing ? (other > key) : (other < key);
          rk += (before || (other == key && j < tid)) ? 1 : 0;
        }
      }
      rank_of[tid] = rk;
      if (rank == 0) {
        if (soa) evals[(size_t)rk * (size_t)batch + (size_t)b] = my_ev;
        else evals[(size_t)b * n + rk] = my_ev;
      }
    }
    if (rank == 0 && tid == 0 && n_iters) {
      int total = 0;
#pragma unroll
      for (unsigned k = 0; k < 4; ++k) total += ld_cluster_i32(cluster_map(smem_u32(rot_total), k));
      n_iters[b] = converged ? total + 1 : 0;
    }
    __syncthreads();
    // ---- eigenvectors straight from registers to their sorted rows
    if (EVECS && v < n) {
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const int pos = 32 * g + k - first;
        if (pos >= 0 && pos < n) {
          const size_t e_idx = (size_t)rank_of[pos] * n + v;
          if (soa) evecs[e_idx * (size_t)batch + (size_t)b] = e[k];
          else evecs[(size_t)b * n * n + e_idx] = e[k];
        }
      }
    }
    cluster_barrier();   // M / ev / counters are reused by the next matrix of this cluster
  }
}

#endif  // __CUDACC__

}  // namespace jpd
