// jpd_cta.cuh -- tier 3: one thread block per matrix, 33 <= n <= 128, the tier-2 scheme
// (jpd_warp.cuh) spread over NPB * NPB/32 threads.
//
// Path being replaced: Jacobi::Diagonalize, /root/reference/include/jacobi_pd.hpp:159-220
// (CalcRot :233-256, ApplyRot :327-393, ApplyRotLeft :409-419, SortRows :477-508).
//
//  * Same odd-even parallel ordering on positions with the exchange folded into the
//    rotation, same even/odd split, conflict-free shared-memory layout of the upper
//    triangle of M, same per-sweep convergence test on all pairs (see jpd_warp.cuh).
//  * E (n x n) lives in REGISTERS: thread (v, ch) owns rows 32 ch .. 32 ch + 31 of column v
//    (n = 128: 512 threads x 32 doubles = the 128 KB of E; with M's 134 KB in shared memory
//    one matrix occupies one SM).  Pattern A pairs never cross a 32-row chunk; in pattern B
//    the pairs (32 ch - 1, 32 ch) do, and the two owners swap that one element through a
//    small shared buffer written before the step's first barrier.
//  * Persistent grid: block b takes matrices b, b + gridDim.x, ... ; the two schedule
//    tables are built once per block.
#pragma once
#include "jpd_common.cuh"
#include "jpd_warp.cuh"

namespace jpd {

template <int NPB> struct CtaShape {
  static constexpr int kChunks = NPB / 32;
  static constexpr int kThreads = NPB * kChunks;
  static constexpr int kH = NPB / 2;
  static constexpr int kLd = NPB + 3;                 // 67, 99, 131: == 3 (mod 8)
  static constexpr int kBankMul = (2 * kLd) & 15;     // 6
  static constexpr int kMElems = NPB * kLd;
  static constexpr int kPairsMax = NPB / 2;
  static constexpr int kHalfWarps = kThreads / 16;
  static constexpr int kIt = 4;                       // block iterations per step (both patterns)
  static constexpr int kTblEntries = kIt * kThreads;
  // elements of T: M | (c,s) interleaved | c | s | eigenvalues | chunk-boundary rows (lo, hi)
  static constexpr int kElems = kMElems + 2 * 4 * kPairsMax + NPB + 2 * kChunks * NPB;   // (c,s) arrays x2 (by pattern)
  JPD_CX static constexpr size_t table_bytes() { return (size_t)2 * kTblEntries * sizeof(unsigned); }
  JPD_CX static constexpr size_t block_bytes(size_t elem) {
    return table_bytes() + ((size_t)kElems * elem + 15) / 16 * 16 + (size_t)(NPB + 4) * sizeof(int);
  }
  JPD_HD static int addr(int r, int c) { return r * kLd + (c >> 1) + (c & 1) * kH; }
};

#if defined(__CUDACC__)

// Same table format as build_schedule() of jpd_warp.cuh: entry [it*threads + tid], the
// handling thread has tid % 16 == bank class of the block.
template <int NPB>
__device__ __forceinline__ void cta_build_schedule(unsigned* tbl, int npr, int off) {
  using Sh = CtaShape<NPB>;
  for (int i = threadIdx.x; i < Sh::kTblEntries; i += Sh::kThreads) tbl[i] = 0xffffffffu;
  __syncthreads();
  if (threadIdx.x < 16) {
    const int cl = threadIdx.x;
    int slot = 0;
    for (int K = 0; K < npr; ++K) {
      for (int L = (cl - Sh::kBankMul * K) & 15; L < npr; L += 16) {
        if (L > K) {
          const int idx = (slot / Sh::kHalfWarps) * Sh::kThreads + (slot % Sh::kHalfWarps) * 16 + cl;
          if (idx < Sh::kTblEntries)
            tbl[idx] = (unsigned)Sh::addr(2 * K + off, 2 * L + off) | ((unsigned)K << 16) | ((unsigned)L << 24);
          ++slot;
        }
      }
    }
  }
}

template <typename T> struct CtaCtx {
  T* M; T* cs; T* cc; T* ss; T* xlo; T* xhi;     // cs/cc/ss: two buffers each, indexed by pattern
  const unsigned* tblA; const unsigned* tblB;
  int m, v, ch;
};

// ---- phase 1: one thread per pair (skip test :191, CalcRot :233-256, own 2x2 block)
template <typename T, int NPB, int OFF>
__device__ __forceinline__ void cta_pivot(const CtaCtx<T>& cx, int& n_rot) {
  using Sh = CtaShape<NPB>;
  using T2 = typename Pair2<T>::type;
  constexpr int o01 = (OFF == 0) ? Sh::kH : 1 - Sh::kH;
  constexpr int o11 = Sh::kLd + o01;
  const int tid = threadIdx.x;
  const int npr = (cx.m >> 1) - OFF;
  if (tid < npr) {
    const int p = 2 * tid + OFF;
    T* pp = cx.M + Sh::addr(p, p);
    T* pq = pp + o01;
    T* qq = pp + o11;
    const T mpp = *pp, mqq = *qq, a = *pq;
    T c, s, npp, nqq;
    if (pivot_exchange_rotation(mpp, mqq, a, c, s, npp, nqq)) {
      *pq = T(0);
      ++n_rot;
    }
    *pp = npp;
    *qq = nqq;
    T2 w; w.x = c; w.y = s;
    reinterpret_cast<T2*>(cx.cs + OFF * 2 * Sh::kPairsMax)[tid] = w;
    cx.cc[OFF * Sh::kPairsMax + tid] = c;
    cx.ss[OFF * Sh::kPairsMax + tid] = s;
  }
}

// ---- phase 2a: 2x2 blocks, B <- G_K B G_L^T (ApplyRot :350-390, positions exchanged), and in
// pattern B the strips of the two unpaired end positions
template <typename T, int NPB, int OFF>
__device__ __forceinline__ void cta_blocks(const CtaCtx<T>& cx) {
  using Sh = CtaShape<NPB>;
  constexpr int o01 = (OFF == 0) ? Sh::kH : 1 - Sh::kH;
  constexpr int o10 = Sh::kLd;
  constexpr int o11 = Sh::kLd + o01;
  const int tid = threadIdx.x;
  const int npr = (cx.m >> 1) - OFF;
  T* M = cx.M;
  const T* cc = cx.cc + OFF * Sh::kPairsMax;
  const T* ss = cx.ss + OFF * Sh::kPairsMax;
  const unsigned* tbl = OFF == 0 ? cx.tblA : cx.tblB;
#pragma unroll
  for (int it = 0; it < Sh::kIt; ++it) {
    const unsigned ent = tbl[it * Sh::kThreads + tid];
    if (ent != 0xffffffffu) {
      T* b = M + (ent & 0xffffu);
      const int K = (ent >> 16) & 0xffu, L = ent >> 24;
      const T ck = cc[K], sk = ss[K], cl = cc[L], sl = ss[L];
      T b00 = b[0], b01 = b[o01], b10 = b[o10], b11 = b[o11];
      exchange_rotate_block(b00, b01, b10, b11, ck, sk, cl, sl);
      b[0] = b00; b[o01] = b01; b[o10] = b10; b[o11] = b11;
    }
  }
  if (OFF == 1 && tid < NPB) {
    const int j = tid < NPB / 2 ? tid : tid - NPB / 2;
    if (j < npr) {
      T* x0 = (tid < NPB / 2) ? M + Sh::addr(0, 2 * j + 1) : M + Sh::addr(2 * j + 1, cx.m - 1);
      T* x1 = (tid < NPB / 2) ? x0 + o01 : x0 + o10;
      T v0 = *x0, v1 = *x1;
      exchange_rotate_pair(v0, v1, cc[j], ss[j]);
      *x0 = v0; *x1 = v1;
    }
  }
}

// ---- phase 2b: eigenvector rows (ApplyRotLeft :414-418) of a step of pattern OFF; registers,
// (c,s) of that step, and for pattern B the two rows received from the neighbouring chunks
template <typename T, int NPB, int OFF>
__device__ __forceinline__ void cta_evecs(const CtaCtx<T>& cx, T (&e)[32]) {
  using Sh = CtaShape<NPB>;
  using T2 = typename Pair2<T>::type;
  const int npr = (cx.m >> 1) - OFF;
  const int ch = cx.ch, v = cx.v;
  const T2* cs2 = reinterpret_cast<const T2*>(cx.cs + OFF * 2 * Sh::kPairsMax);
  const int j0 = 16 * ch;   // pair index of this chunk's first local pair
  T new_lo = e[0], new_hi = e[31];
  if (OFF == 1) {
    // pairs that cross the chunk boundary: (32ch-1, 32ch) = pair j0-1, (32ch+31, 32ch+32) = pair j0+15
    if (ch > 0 && j0 - 1 < npr) {
      const T2 r = cs2[j0 - 1];
      new_lo = jfma(r.x, cx.xhi[(ch - 1) * NPB + v], -(r.y * e[0]));
    }
    if (ch < Sh::kChunks - 1 && j0 + 15 < npr) {
      const T2 r = cs2[j0 + 15];
      new_hi = jfma(r.y, e[31], r.x * cx.xlo[(ch + 1) * NPB + v]);
    }
  }
#pragma unroll
  for (int jj = 0; jj < 16 - OFF; ++jj) {
    const int p = 2 * jj + OFF, q = p + 1;
    if (j0 + jj < npr) {
      const T2 r = cs2[j0 + jj];
      exchange_rotate_pair(e[p], e[q], r.x, r.y);
    }
  }
  if (OFF == 1) { e[0] = new_lo; e[31] = new_hi; }
}

// One A step + one B step.  Phase 1 occupies npr <= 64 threads while the block waits, so the
// eigenvector phase of a step is deferred to the next step and runs beside its phase 1:
//   A: pivot_A, E(previous B) | barrier | blocks_A | barrier
//   B: pivot_B, E(this A), publish the chunk-boundary rows | barrier | blocks_B | barrier
// ((c,s) double-buffered by pattern: E reads the other pattern's buffer than pivot writes.)
template <typename T, int NPB, bool EVECS>
__device__ __forceinline__ void cta_step_pair(const CtaCtx<T>& cx, T (&e)[EVECS ? 32 : 1], bool have_prev,
                                              int& n_rot) {
  using Sh = CtaShape<NPB>;
  cta_pivot<T, NPB, 0>(cx, n_rot);
  if constexpr (EVECS) { if (have_prev) cta_evecs<T, NPB, 1>(cx, e); }
  __syncthreads();
  cta_blocks<T, NPB, 0>(cx);
  __syncthreads();

  cta_pivot<T, NPB, 1>(cx, n_rot);
  if constexpr (EVECS) {
    cta_evecs<T, NPB, 0>(cx, e);
    if (cx.ch > 0) cx.xlo[cx.ch * NPB + cx.v] = e[0];
    if (cx.ch < Sh::kChunks - 1) cx.xhi[cx.ch * NPB + cx.v] = e[31];
  }
  __syncthreads();
  cta_blocks<T, NPB, 1>(cx);
  __syncthreads();
}

template <typename T, int NPB, bool EVECS>
__global__ void __launch_bounds__(CtaShape<NPB>::kThreads, (NPB == 64 ? 4 : (NPB == 96 ? 2 : 1)))
jpd_cta_kernel(const T* __restrict__ mat, T* __restrict__ evals, T* __restrict__ evecs,
               int* __restrict__ n_iters, int n, long long batch, int soa, int sort_criteria,
               int max_num_sweeps) {
  using Sh = CtaShape<NPB>;
  // internal callers (tier 5's capped pivot solves) pass -(cap): an unfinished solve then reports -(rotations + 1)
  const bool raw_count = max_num_sweeps < 0;
  if (raw_count) max_num_sweeps = -max_num_sweeps;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int v = tid % NPB, ch = tid / NPB;
  const int m = n + (n & 1);

  unsigned* tblA = reinterpret_cast<unsigned*>(smem_raw);
  unsigned* tblB = tblA + Sh::kTblEntries;
  cta_build_schedule<NPB>(tblA, m >> 1, 0);
  cta_build_schedule<NPB>(tblB, (m >> 1) - 1, 1);

  T* M = reinterpret_cast<T*>(smem_raw + Sh::table_bytes());
  T* cs = M + Sh::kMElems;
  T* cc = cs + 2 * 2 * Sh::kPairsMax;
  T* ss = cc + 2 * Sh::kPairsMax;
  T* ev = ss + 2 * Sh::kPairsMax;
  T* xlo = ev + NPB;
  T* xhi = xlo + Sh::kChunks * NPB;
  int* perm = reinterpret_cast<int*>(smem_raw + Sh::table_bytes() + ((size_t)Sh::kElems * sizeof(T) + 15) / 16 * 16);
  int* rot_total = perm + NPB;
  CtaCtx<T> cx;
  cx.M = M; cx.cs = cs; cx.cc = cc; cx.ss = ss; cx.xlo = xlo; cx.xhi = xhi;
  cx.tblA = tblA; cx.tblB = tblB; cx.m = m; cx.v = v; cx.ch = ch;
  __syncthreads();

  for (long long b = blockIdx.x; b < batch; b += gridDim.x) {
    // ---- load: upper triangle only (jacobi_pd.hpp:167-171); everything else zero
    for (int i = tid; i < Sh::kMElems; i += Sh::kThreads) M[i] = T(0);
    if (tid == 0) *rot_total = 0;
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
      for (int k = 0; k < 32; ++k) e[k] = (32 * ch + k == v) ? T(1) : T(0);   // :172-178
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
      for (int step = 0; step < m; step += 2) cta_step_pair<T, NPB, EVECS>(cx, e, step > 0, n_rot);
      if constexpr (EVECS) cta_evecs<T, NPB, 1>(cx, e);   // the last B step's eigenvector phase
      ++sweeps;
    }
    // max_num_sweeps == 0: the reference reports failure for n > 1 whatever the matrix (:185-186, :214-219)
    const bool converged = !need && max_num_sweeps > 0;
    if (n_rot) atomicAdd(rot_total, n_rot);

    // ---- eigenvalues (:207-209); a sweep reverses the order of the positions
    const int first = ((n & 1) && (sweeps & 1)) ? 1 : 0;
    T my_ev = T(0);
    if (tid < n) { my_ev = M[Sh::addr(first + tid, first + tid)]; ev[tid] = my_ev; }
    __syncthreads();
    if (tid < n) {
      int rank = tid;
      if (sort_criteria >= SORT_DECREASING_EVALS && sort_criteria <= SORT_INCREASING_ABS_EVALS) {
        const bool use_abs = sort_criteria >= SORT_DECREASING_ABS_EVALS;
        const bool decreasing = (sort_criteria == SORT_DECREASING_EVALS) || (sort_criteria == SORT_DECREASING_ABS_EVALS);
        const long long key = sortable_bits(use_abs ? jabs(my_ev) : my_ev);
        rank = 0;
        for (int j = 0; j < n; ++j) {
          const T o = ev[j];
          const long long other = sortable_bits(use_abs ? jabs(o) : o);
          const bool before = decreasing ? (other > key) : (other < key);
          rank += (before || (other == key && j < tid)) ? 1 : 0;
        }
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
      for (int k = 0; k < 32; ++k) Es[(32 * ch + k) * NPB + v] = e[k];
      __syncthreads();
      for (int idx = tid; idx < n * n; idx += Sh::kThreads) {
        const int k = idx / n, c = idx - k * n;
        const T val = Es[(first + perm[k]) * NPB + c];
        if (soa) evecs[(size_t)idx * (size_t)batch + (size_t)b] = val;
        else evecs[(size_t)b * n * n + idx] = val;
      }
    }
    __syncthreads();   // M / ev / perm are reused by the next matrix of this block
  }
}

#endif  // __CUDACC__

}  // namespace jpd
