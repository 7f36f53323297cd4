// jpd_blockjacobi.cuh -- tier 5: block Jacobi for 257 <= n <= 1024 (the reference benchmarks
// its header up to n = 1000, /root/reference/benchmarks/README.md:19-20).
//
// Path being replaced: Jacobi::Diagonalize, /root/reference/include/jacobi_pd.hpp:159-220.
//
// A matrix of this size is worked on by the WHOLE device, one outer step at a time:
//  * The indices are cut into blocks of 32 (the matrix is padded to a multiple of 64 with
//    decoupled zero rows/columns, which no rotation ever touches: their off-diagonal entries
//    are exact zeros).  An outer sweep is the round-robin tournament of the block indices:
//    nb-1 steps, each pairing all blocks into nb/2 disjoint pairs (I,J).
//  * For every pair the 64 x 64 pivot matrix M[I u J, I u J] is gathered (blk_gather) and
//    diagonalised by the tier-3 kernel (jpd_cta.cuh: the odd-even two-sided Jacobi with the
//    reference's skip test :191 and CalcRot :233-256) -- a batch of nb/2 x matrices
//    independent 64 x 64 problems; its eigenvector rows, SORTED by eigenvalue so that Q tends
//    to the identity as the matrix converges (an arbitrary row order makes block Jacobi
//    converge only linearly), are the orthogonal block Q.
//  * ApplyRot / ApplyRotLeft (:327-393, :409-419) for ALL the rotations of those solves at once
//    are two small GEMMs per 64 x 64 tile: M[R_k, R_l] <- Q_k M[R_k, R_l] Q_l^T and
//    E[R_k, :] <- Q_k E[R_k, :] (blk_apply), register-tiled FP64 FMA.  A product of the ~4000
//    rotations of one pivot solve applied as a dense 64 x 64 block costs 64 FMAs per entry
//    instead of ~250, and every entry of M makes ONE round trip through shared memory per
//    outer step instead of one per rotation step.
//  * Convergence: every index pair belongs to exactly one pivot block per sweep, so a sweep
//    in which no pivot solve performed a rotation is a sweep in which every pair passed the
//    reference's test -- the same stopping rule as the other tiers.  The host reads one flag
//    per sweep (the call therefore synchronises its stream once per outer sweep).
#pragma once
#include <type_traits>
#include "jpd_common.cuh"

namespace jpd {

constexpr int kBjBlock = 32;          // indices per block
constexpr int kBjPivot = 64;          // pivot matrix = two blocks
constexpr int kBjLd = 68;             // shared-memory row pitch (elements): rows 16-byte aligned in fp32 and fp64
#ifndef JPD_BJ_APPLY_BUFFERS
#define JPD_BJ_APPLY_BUFFERS 2   // measured: 2 (three resident blocks per SM) beats 3 (Q_l^T prefetched) by 3-5 %, tests/scratch/t5vv.py
#endif
constexpr int kBjThreads = 256;       // gather kernel
constexpr int kBjTR = 8;              // output rows per thread of the 64 x 64 x 64 products (x 4 columns)
constexpr int kBjGemmThreads = 16 * (kBjPivot / kBjTR);   // 128

#if defined(__CUDACC__)

// ---- workspace initialisation: symmetric padded M from the caller's upper triangle (:167-171), E = I
template <typename T>
__global__ void __launch_bounds__(256)
blk_init_kernel(const T* __restrict__ mat, T* __restrict__ Mw, T* __restrict__ Ew, int n, int np,
                long long batch, long long b0, long long count, int soa) {
  const long long total = count * (long long)np * np;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long lb = idx / ((long long)np * np);
    const int rem = (int)(idx - lb * (long long)np * np);
    const int r = rem / np, c = rem - r * np;
    T v = T(0);
    if (r < n && c < n) {
      const int i = r < c ? r : c, j = r < c ? c : r;
      const size_t e = (size_t)i * n + j;
      const long long b = b0 + lb;
      v = soa ? mat[e * (size_t)batch + (size_t)b] : mat[(size_t)b * n * n + e];
    }
    Mw[idx] = v;
    Ew[idx] = (r == c) ? T(1) : T(0);
  }
}

// rows of pair (I,J): local row q (0..63) -> global row
__device__ __forceinline__ int bj_row(int I, int J, int q) {
  return (q < kBjBlock ? I * kBjBlock : J * kBjBlock - kBjBlock) + q;
}

// ---- pivot matrices of one step: P[(lb * npairs + k)] = M[R_k, R_k], AoS 64 x 64
template <typename T>
__global__ void __launch_bounds__(kBjThreads)
blk_gather_kernel(const T* __restrict__ Mw, T* __restrict__ P, const int2* __restrict__ pairs, int npairs, int np) {
  const int k = blockIdx.x, lb = blockIdx.y;
  const int2 pr = pairs[k];
  const T* M = Mw + (size_t)lb * np * np;
  T* out = P + ((size_t)lb * npairs + k) * (kBjPivot * kBjPivot);
  for (int idx = threadIdx.x; idx < kBjPivot * kBjPivot; idx += kBjThreads) {
    const int r = idx >> 6, c = idx & 63;
    out[idx] = M[(size_t)bj_row(pr.x, pr.y, r) * np + bj_row(pr.x, pr.y, c)];
  }
}

// 16-byte asynchronous copies global -> shared (LDGSTS): a 64-element row in 64 * sizeof(T) / 16 pieces
__device__ __forceinline__ void bj_cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void bj_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bj_cp_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// a 64 x 64 block with row pitch `src_pitch` whose 64 columns are two runs of 32 (at col0 and col1) -> shared memory,
// row-major with pitch kBjLd; row r of the block is source row rowmap(r)
template <typename T, typename RowMap>
__device__ __forceinline__ void bj_load_block_async(T* __restrict__ dst, const T* __restrict__ src, size_t src_pitch,
                                                    int col0, int col1, RowMap rowmap) {
  constexpr int kPer = 16 / (int)sizeof(T);             // elements per piece
  constexpr int kPieces = kBjPivot / kPer;              // pieces per row
  for (int idx = threadIdx.x; idx < kBjPivot * kPieces; idx += kBjGemmThreads) {
    const int r = idx / kPieces, c = (idx - r * kPieces) * kPer;
    const int gc = c < kBjBlock ? col0 + c : col1 + (c - kBjBlock);
    bj_cp_async16(dst + r * kBjLd + c, src + (size_t)rowmap(r) * src_pitch + gc);
  }
}

// C += A B for one 64 x 64 x 64 product, 8 x 4 outputs per thread (128 threads): rows 8ty .. 8ty+7, columns tx, tx+16,
// tx+32, tx+48.  At = A TRANSPOSED in shared memory (At[l][row]), B row-major (B[l][col]), pitch kBjLd.  Per l a
// thread fetches its eight a's as four 16-byte vectors (the same for 16 lanes) and its four b's from four runs of 16
// consecutive elements: 16 shared-memory wavefronts per 32 FMAs.  History (n = 512, application kernel, ncu):
// 4 x 4 outputs with a's from four rows and b's at a stride of four elements: 20 wavefronts per 16 FMAs, 40 % of them
// bank conflicts, FP64 pipe 23 % busy; 4 x 4 with this layout: 12 per 16, FP64 51 %, shared pipe 81 %.
template <typename T>
__device__ __forceinline__ void bj_gemm_64(const T* __restrict__ At, const T* __restrict__ B, int ty, int tx, T (&c)[kBjTR][4]) {
  using T2 = typename std::conditional<sizeof(T) == 8, double2, float2>::type;
#pragma unroll 8
  for (int l = 0; l < kBjPivot; ++l) {
    T a[kBjTR], b[4];
#pragma unroll
    for (int i = 0; i < kBjTR; i += 2) {
      const T2 v = *reinterpret_cast<const T2*>(At + l * kBjLd + kBjTR * ty + i);
      a[i] = v.x; a[i + 1] = v.y;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = B[l * kBjLd + tx + 16 * j];
#pragma unroll
    for (int i = 0; i < kBjTR; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) c[i][j] = jfma(a[i], b[j], c[i][j]);
  }
}

// ---- One Newton-Schulz step  Q <- Q + (I - Q Q^T) Q / 2  on every pivot solve's Q.  The inner
// solver accumulates ~10^4 rotations per 64 x 64 block, which leaves Q orthogonal only to
// ~100 ulp; applied some hundreds of times per matrix (steps x sweeps) that drift would end
// up in the eigenvalues (measured without this step: 1.2e-12 max|lambda| at n = 1000 in fp64,
// 5e-5 at n = 257 in fp32 -- at and beyond the north-star tolerance).  One step squares the
// defect; it costs two 64^3 products per pivot, a few percent of the application.
template <typename T>
__global__ void __launch_bounds__(kBjGemmThreads, 2)
blk_orthonormalize_kernel(T* __restrict__ Q) {
  extern __shared__ __align__(16) unsigned char bj_smem_raw[];
  T* sQ = reinterpret_cast<T*>(bj_smem_raw);        // Q row-major: B operand of R Q
  T* sRt = sQ + kBjPivot * kBjLd;                   // R = (I - Q Q^T) / 2, transposed (A operand)
  T* sQt = sRt + kBjPivot * kBjLd;                  // Q^T row-major = Q transposed: A and B operand of Q Q^T
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  T* q = Q + (size_t)blockIdx.x * (kBjPivot * kBjPivot);
  for (int idx = tid; idx < kBjPivot * kBjPivot; idx += kBjGemmThreads) {
    const int r = idx >> 6, c = idx & 63;
    const T v = q[idx];
    sQ[r * kBjLd + c] = v;
    sQt[c * kBjLd + r] = v;
  }
  __syncthreads();
  T acc[kBjTR][4];
#pragma unroll
  for (int i = 0; i < kBjTR; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
  bj_gemm_64(sQt, sQt, ty, tx, acc);                // G = Q Q^T
#pragma unroll
  for (int i = 0; i < kBjTR; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = kBjTR * ty + i, c = tx + 16 * j;
      sRt[c * kBjLd + r] = T(0.5) * ((r == c ? T(1) : T(0)) - acc[i][j]);
      acc[i][j] = sQ[r * kBjLd + c];                // Q + R Q
    }
  __syncthreads();
  bj_gemm_64(sRt, sQ, ty, tx, acc);
  // the result leaves TRANSPOSED (blk_apply wants Q_k^T as the A operand of Q_k X and Q_l^T as the B operand of
  // (Q_k X) Q_l^T: both become plain row-major copies there), staged through sQt for coalesced stores
#pragma unroll
  for (int i = 0; i < kBjTR; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) sQt[(tx + 16 * j) * kBjLd + kBjTR * ty + i] = acc[i][j];      // free since the first product
  __syncthreads();
  for (int idx = tid; idx < kBjPivot * kBjPivot; idx += kBjGemmThreads) q[idx] = sQt[(idx >> 6) * kBjLd + (idx & 63)];
}

// ---- ApplyRot + ApplyRotLeft of one outer step.  Work items (blockIdx.x) per matrix:
//   [0, ntile)          tile (k <= l) of M:  M[R_k, R_l] <- Q_k M[R_k, R_l] Q_l^T, and its mirror image
//   [ntile, ntile + npairs * ncol)   (k, col block c) of E:  E[R_k, C_c] <- Q_k E[R_k, C_c]
// Q_k = eigenvector ROWS of pivot solve k (row a = new index a); Qt holds the blocks TRANSPOSED
// (blk_orthonormalize_kernel), so both operands that involve Q are plain row-major copies.  All three operands
// arrive as 16-byte asynchronous copies; Q_l^T, which only the second product needs, lands during the first.
template <typename T, bool EVECS>
__global__ void __launch_bounds__(kBjGemmThreads, JPD_BJ_APPLY_BUFFERS == 2 ? 3 : 2)
blk_apply_kernel(T* __restrict__ Mw, T* __restrict__ Ew, const T* __restrict__ Qt,
                 const int2* __restrict__ pairs, int npairs, int np) {
  extern __shared__ __align__(16) unsigned char bj_smem_raw[];
  T* sQat = reinterpret_cast<T*>(bj_smem_raw);     // Q_k transposed (A operand)
  T* sX = sQat + kBjPivot * kBjLd;                   // the tile (B operand), later T = Q_k X transposed (A operand)
  // Q_l^T (B operand of the second product): a third buffer that fills during the first product, or -- two buffers,
  // three resident blocks per SM instead of two -- Q_k^T's buffer once the first product is done
  T* sQbt = JPD_BJ_APPLY_BUFFERS == 2 ? sQat : sX + kBjPivot * kBjLd;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int lb = blockIdx.y;
  const int ntile = npairs * (npairs + 1) / 2;
  const int item = blockIdx.x;
  T* M = Mw + (size_t)lb * np * np;
  const T* Qb = Qt + (size_t)lb * npairs * (kBjPivot * kBjPivot);
  auto same_row = [](int r) { return r; };

  if (item < ntile) {
    // (k, l), k <= l, from the linear index
    int k = 0, rest = item;
    while (rest >= npairs - k) { rest -= npairs - k; ++k; }
    const int l = k + rest;
    const int2 pk = pairs[k], pl = pairs[l];
    bj_load_block_async(sQat, Qb + (size_t)k * 4096, kBjPivot, 0, kBjBlock, same_row);
    bj_load_block_async(sX, M, (size_t)np, bj_row(pl.x, pl.y, 0), bj_row(pl.x, pl.y, kBjBlock),
                        [&](int r) { return bj_row(pk.x, pk.y, r); });
    bj_cp_commit();
    if (JPD_BJ_APPLY_BUFFERS != 2) bj_load_block_async(sQbt, Qb + (size_t)l * 4096, kBjPivot, 0, kBjBlock, same_row);
    bj_cp_commit();
    bj_cp_wait<1>();                                   // Q_k^T and the tile are here
    __syncthreads();
    T acc[kBjTR][4];
#pragma unroll
    for (int i = 0; i < kBjTR; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
    bj_gemm_64(sQat, sX, ty, tx, acc);                // T = Q_k X
    __syncthreads();                                   // every thread has read X (and Q_k^T)
    if (JPD_BJ_APPLY_BUFFERS == 2) { bj_load_block_async(sQbt, Qb + (size_t)l * 4096, kBjPivot, 0, kBjBlock, same_row); bj_cp_commit(); }
#pragma unroll
    for (int i = 0; i < kBjTR; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { sX[(tx + 16 * j) * kBjLd + kBjTR * ty + i] = acc[i][j]; acc[i][j] = T(0); }   // T transposed
    bj_cp_wait<0>();                                   // Q_l^T
    __syncthreads();
    bj_gemm_64(sX, sQbt, ty, tx, acc);                // T Q_l^T
#pragma unroll
    for (int i = 0; i < kBjTR; ++i) {
      const int gr = bj_row(pk.x, pk.y, kBjTR * ty + i);
#pragma unroll
      for (int j = 0; j < 4; ++j) M[(size_t)gr * np + bj_row(pl.x, pl.y, tx + 16 * j)] = acc[i][j];
    }
    if (k != l) {
      // the mirror image (the matrix stays symmetric bit for bit): transposed through shared memory -- sQat is free
      // since the first product -- so that it leaves in 256-byte runs instead of 4096 scattered 8-byte stores
      if (JPD_BJ_APPLY_BUFFERS == 2) __syncthreads();  // ... with two buffers: since the second product
#pragma unroll
      for (int i = 0; i < kBjTR; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) sQat[(tx + 16 * j) * kBjLd + kBjTR * ty + i] = acc[i][j];
      __syncthreads();
      for (int idx = tid; idx < kBjPivot * kBjPivot; idx += kBjGemmThreads) {
        const int r = idx >> 6, c = idx & 63;
        M[(size_t)bj_row(pl.x, pl.y, r) * np + bj_row(pk.x, pk.y, c)] = sQat[r * kBjLd + c];
      }
    }
    if (k == l) {
      // symmetrise the diagonal tile's two triangles (Q P Q^T is symmetric only up to rounding)
      __syncthreads();
      for (int idx = tid; idx < kBjPivot * kBjPivot; idx += kBjGemmThreads) {
        const int r = idx >> 6, c = idx & 63;
        if (r > c) {
          const int gr = bj_row(pk.x, pk.y, r), gc = bj_row(pk.x, pk.y, c);
          M[(size_t)gr * np + gc] = M[(size_t)gc * np + gr];
        }
      }
    }
  } else if (EVECS) {
    const int e_item = item - ntile;
    const int ncol = np / kBjPivot;
    const int k = e_item / ncol, cb = e_item - k * ncol;
    const int2 pk = pairs[k];
    T* E = Ew + (size_t)lb * np * np;
    bj_load_block_async(sQat, Qb + (size_t)k * 4096, kBjPivot, 0, kBjBlock, same_row);
    bj_load_block_async(sX, E, (size_t)np, cb * kBjPivot, cb * kBjPivot + kBjBlock,
                        [&](int r) { return bj_row(pk.x, pk.y, r); });
    bj_cp_commit();
    bj_cp_wait<0>();
    __syncthreads();
    T acc[kBjTR][4];
#pragma unroll
    for (int i = 0; i < kBjTR; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
    bj_gemm_64(sQat, sX, ty, tx, acc);
#pragma unroll
    for (int i = 0; i < kBjTR; ++i) {
      const int gr = bj_row(pk.x, pk.y, kBjTR * ty + i);
#pragma unroll
      for (int j = 0; j < 4; ++j) E[(size_t)gr * np + cb * kBjPivot + tx + 16 * j] = acc[i][j];
    }
  }
}

// ---- bookkeeping of one step: rotations of the pivot solves, per matrix.  piv_iters follows the
// reference's return value (rotations + 1, or 0 when the inner solve ran out of sweeps).
__global__ void blk_count_kernel(const int* __restrict__ piv_iters, int npairs, long long count,
                                 int* __restrict__ rot_total, int* __restrict__ rot_sweep, int* __restrict__ any_left) {
  const long long lb = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (lb >= count) return;
  int rot = 0;
  for (int k = 0; k < npairs; ++k) {
    const int it = piv_iters[lb * npairs + k];
    rot += it > 0 ? it - 1 : (it < 0 ? -it - 1 : 1 << 16);   // < 0: a capped inner solve, -(rotations + 1): it certainly rotated
  }
  if (rot) {
    rot_total[lb] = min(rot_total[lb] + rot, 0x3fffffff);
    rot_sweep[lb] += 1;
    *any_left = 1;
  }
}

// called between sweeps: remembers which matrices were already done BEFORE the sweep that just ended
__global__ void blk_sweep_end_kernel(int* __restrict__ rot_sweep, int* __restrict__ converged, long long count) {
  const long long lb = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (lb >= count) return;
  if (rot_sweep[lb] == 0) converged[lb] = 1;     // a whole sweep without a rotation (:191 on every pair)
  rot_sweep[lb] = 0;
}

// ---- eigenvalues (:207-209), order (SortRows :477-508: the keys are ranked, ties by index as in
// tiers 2-4) and output.  Rows that belong to the padding are exact unit vectors on a padding
// column and are left out.  One block per matrix.
template <typename T, bool EVECS>
__global__ void __launch_bounds__(256)
blk_finalize_kernel(const T* __restrict__ Mw, const T* __restrict__ Ew, const int* __restrict__ rot_total,
                    const int* __restrict__ converged, T* __restrict__ evals, T* __restrict__ evecs,
                    int* __restrict__ n_iters, int n, int np, long long batch, long long b0, int soa,
                    int sort_criteria, int max_num_sweeps) {
  extern __shared__ __align__(16) unsigned char bj_smem_raw[];
  T* key = reinterpret_cast<T*>(bj_smem_raw);            // [np] diagonal of M
  int* is_pad = reinterpret_cast<int*>(key + np);        // [np] row belongs to the padding
  int* pos_of = is_pad + np;                             // [np] output position of row r
  const int lb = blockIdx.x;
  const long long b = b0 + lb;
  const T* M = Mw + (size_t)lb * np * np;
  const T* E = Ew + (size_t)lb * np * np;
  for (int r = threadIdx.x; r < np; r += blockDim.x) {
    bool pad = false;
    for (int c = n; c < np; ++c) pad = pad || !is_zero(E[(size_t)r * np + c]);
    key[r] = M[(size_t)r * np + r];
    is_pad[r] = pad ? 1 : 0;
  }
  __syncthreads();
  const bool sorted = sort_criteria >= SORT_DECREASING_EVALS && sort_criteria <= SORT_INCREASING_ABS_EVALS;
  const bool use_abs = sort_criteria >= SORT_DECREASING_ABS_EVALS;
  const bool decreasing = (sort_criteria == SORT_DECREASING_EVALS) || (sort_criteria == SORT_DECREASING_ABS_EVALS);
  for (int r = threadIdx.x; r < np; r += blockDim.x) {
    int rank = np;                                        // padding rows: beyond every output row
    if (!is_pad[r]) {
      const T mine = use_abs ? jabs(key[r]) : key[r];
      rank = 0;
      for (int j = 0; j < np; ++j) {
        if (j == r || is_pad[j]) continue;
        const T other = use_abs ? jabs(key[j]) : key[j];
        const bool strictly = sorted && (decreasing ? (other > mine) : (other < mine));
        const bool tie = !sorted || other == mine;
        rank += (strictly || (tie && j < r)) ? 1 : 0;
      }
    }
    pos_of[r] = rank;
  }
  __syncthreads();
  for (int r = threadIdx.x; r < np; r += blockDim.x) {
    const int pos = pos_of[r];
    if (pos < n) {
      if (soa) evals[(size_t)pos * (size_t)batch + (size_t)b] = key[r];
      else evals[(size_t)b * n + pos] = key[r];
    }
  }
  if (threadIdx.x == 0 && n_iters) n_iters[b] = (converged[lb] && max_num_sweeps > 0) ? rot_total[lb] + 1 : 0;
  if (EVECS) {
    for (int r = 0; r < np; ++r) {
      const int pos = pos_of[r];
      if (pos >= n) continue;
      for (int c = threadIdx.x; c < n; c += blockDim.x) {
        const T v = E[(size_t)r * np + c];
        const size_t e_idx = (size_t)pos * n + c;
        if (soa) evecs[e_idx * (size_t)batch + (size_t)b] = v;
        else evecs[(size_t)b * n * n + e_idx] = v;
      }
    }
  }
}

#endif  // __CUDACC__

}  // namespace jpd
