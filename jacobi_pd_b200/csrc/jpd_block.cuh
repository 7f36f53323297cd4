// jpd_block.cuh -- tier 5 (n > 256; the first-generation kernel, which ran n = 7..256 until
// jpd_warp.cuh / jpd_cta.cuh / jpd_cluster.cuh replaced it there): one thread block per matrix,
// round-robin parallel-ordering sweeps.
//
// Path being replaced: Jacobi::Diagonalize, /root/reference/include/jacobi_pd.hpp:159-220.
// The reference applies ONE max-pivot rotation at a time (MaxEntry :444-468); a block
// instead applies the n/2 disjoint rotations of one round-robin (Brent-Luk /
// chess-tournament) step at once:
//
//   phase A   thread k < n/2 owns pair (p_k, q_k): the reference's skip test (:191),
//             the rotation (CalcRot :233-256), and the pair's own 2x2 diagonal block
//             (M_pp -= t a, M_qq += t a, M_pq = 0; :337-338,347).
//   phase B   every 2x2 block {rows p_k,q_k} x {cols p_l,q_l}, k<l, gets R_k^T B R_l
//             (ApplyRot :350-390 for both rotations at once); the mirror image is
//             written too so that M stays a full symmetric matrix and the next step can
//             pair any two indices.  Eigenvector rows p_k,q_k get R_k^T (:414-418).
//
// Storage: M (n x ldm, ldm odd => conflict-free column access) and E (n x n) live in
// shared memory when they fit, otherwise (always, at the sizes that still reach this kernel)
// in a per-block global workspace that stays L2 resident.  Convergence: a full sweep in which no pair needed a
// rotation (block vote through __syncthreads_or).  Return value as the reference:
// rotations + 1, or 0 when max_num_sweeps sweeps were not enough (:214-219).
#pragma once
#include "jpd_common.cuh"

namespace jpd {

struct BlockPlan {
  int threads;        // block size
  int ldm;            // leading dimension of M (odd)
  int m_in_smem;      // 1: M in shared memory, 0: in the global workspace
  int e_in_smem;      // same for E
  size_t smem_bytes;  // dynamic shared memory per block
  size_t work_elems;  // global workspace elements per block (0 if none)
};

#if defined(__CUDACC__)

// circle method: step s, table k of half = m/2 (m even, player m-1 fixed)
__device__ __forceinline__ void rr_pair(int m, int s, int k, int& p, int& q) {
  int a, b;
  if (k == 0) { a = m - 1; b = s; }
  else {
    a = s + k;  if (a >= m - 1) a -= (m - 1);
    b = s - k;  if (b < 0) b += (m - 1);
  }
  p = a < b ? a : b;
  q = a < b ? b : a;
}

template <typename T>
__global__ void jpd_block_kernel(const T* __restrict__ mat, T* __restrict__ evals,
                                 T* __restrict__ evecs, int* __restrict__ n_iters, int n,
                                 long long batch, int soa, int sort_criteria, int calc_evecs,
                                 int max_num_sweeps, int ldm, int m_in_smem, int e_in_smem,
                                 T* __restrict__ gwork, size_t work_elems) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int m = n + (n & 1), half = m >> 1, steps = m - 1;

  // ---- carve shared memory: [cs: half x (c,s)] [key: n] [M?] [E?] [id: n ints] [misc]
  T* sm = reinterpret_cast<T*>(smem_raw);
  T* cs = sm;                       sm += 2 * half;
  T* key = sm;                      sm += n;
  T* Ms; T* Es;
  T* gw = gwork ? gwork + (size_t)blockIdx.x * work_elems : nullptr;
  if (m_in_smem) { Ms = sm; sm += (size_t)n * ldm; } else { Ms = gw; gw += (size_t)n * ldm; }
  if (e_in_smem) { Es = sm; if (calc_evecs) sm += (size_t)n * n; } else { Es = gw; }
  int* ids = reinterpret_cast<int*>(sm);
  int* rot_total = ids + n;

  for (long long b = blockIdx.x; b < batch; b += gridDim.x) {
    // ---- load: upper triangle only (jacobi_pd.hpp:167-171), mirrored into full storage
    for (int idx = tid; idx < n * n; idx += nthr) {
      const int i = idx / n, j = idx - i * n;
      const int r = i < j ? i : j, c = i < j ? j : i;
      const T v = soa ? mat[((size_t)r * n + c) * (size_t)batch + (size_t)b]
                      : mat[(size_t)b * n * n + (size_t)r * n + c];
      Ms[(size_t)i * ldm + j] = v;
      if (calc_evecs) Es[(size_t)i * n + j] = (i == j) ? T(1) : T(0);   // :172-178
    }
    if (tid == 0) *rot_total = 0;
    __syncthreads();

    int my_rot = 0;
    bool converged = (n == 1);
    for (int sweep = 0; sweep < max_num_sweeps && !converged; ++sweep) {
      int sweep_rotated = 0;
      for (int s = 0; s < steps; ++s) {
        // ---------------- phase A: one thread per pair
        int need_i = 0;
        if (tid < half) {
          int p, q;
          rr_pair(m, s, tid, p, q);
          T c = T(1), sn = T(0);
          if (q < n) {  // q == n only for the dummy player of odd n
            const T mpp = Ms[(size_t)p * ldm + p], mqq = Ms[(size_t)q * ldm + q];
            const T a = Ms[(size_t)p * ldm + q];
            if (pair_needs_rotation(a, mpp, mqq)) {
              const T d = mqq - mpp;
              T t;
              jacobi_rotation(d, a, rotation_range_unsafe(d, a), c, sn, t);
              const T g = t * a;
              Ms[(size_t)p * ldm + p] = mpp - g;
              Ms[(size_t)q * ldm + q] = mqq + g;
              Ms[(size_t)p * ldm + q] = T(0);
              Ms[(size_t)q * ldm + p] = T(0);
              need_i = 1;
              ++my_rot;
            }
          }
          cs[2 * tid] = c;
          cs[2 * tid + 1] = sn;
        }
        const int any = __syncthreads_or(need_i);
        sweep_rotated |= any;
        if (!any) continue;  // uniform: nothing to apply in this step

        // ---------------- phase B: off-diagonal 2x2 blocks {k, k+delta}, then E rows
        const int n_delta = (half - 1) / 2;           // full circulant offsets
        const int n_full = n_delta * half;
        const int n_blocks = n_full + ((half & 1) ? 0 : half / 2);
        for (int it = tid; it < n_blocks; it += nthr) {
          int k, l;
          if (it < n_full) { const int dl = it / half; k = it - dl * half; l = k + dl + 1; if (l >= half) l -= half; }
          else { k = it - n_full; l = k + half / 2; }
          int pk, qk, pl, ql;
          rr_pair(m, s, k, pk, qk);
          rr_pair(m, s, l, pl, ql);
          const T ck = cs[2 * k], sk = cs[2 * k + 1], cl = cs[2 * l], sl = cs[2 * l + 1];
          const bool qk_ok = qk < n, ql_ok = ql < n;
          // B = [[b00 b01],[b10 b11]] rows (pk,qk) x cols (pl,ql)
          const T b00 = Ms[(size_t)pk * ldm + pl];
          const T b01 = ql_ok ? Ms[(size_t)pk * ldm + ql] : T(0);
          const T b10 = qk_ok ? Ms[(size_t)qk * ldm + pl] : T(0);
          const T b11 = (qk_ok && ql_ok) ? Ms[(size_t)qk * ldm + ql] : T(0);
          // rows: R_k^T B  (row_p' = c row_p - s row_q ; row_q' = s row_p + c row_q)
          const T r00 = ck * b00 - sk * b10, r01 = ck * b01 - sk * b11;
          const T r10 = sk * b00 + ck * b10, r11 = sk * b01 + ck * b11;
          // cols: (.) R_l (col_p' = c col_p - s col_q ; col_q' = s col_p + c col_q)
          const T n00 = cl * r00 - sl * r01, n01 = sl * r00 + cl * r01;
          const T n10 = cl * r10 - sl * r11, n11 = sl * r10 + cl * r11;
          Ms[(size_t)pk * ldm + pl] = n00;  Ms[(size_t)pl * ldm + pk] = n00;
          if (ql_ok) { Ms[(size_t)pk * ldm + ql] = n01;  Ms[(size_t)ql * ldm + pk] = n01; }
          if (qk_ok) { Ms[(size_t)qk * ldm + pl] = n10;  Ms[(size_t)pl * ldm + qk] = n10; }
          if (qk_ok && ql_ok) { Ms[(size_t)qk * ldm + ql] = n11;  Ms[(size_t)ql * ldm + qk] = n11; }
        }
        if (calc_evecs) {
          for (int it = tid; it < half * n; it += nthr) {
            const int k = it / n, v = it - k * n;
            int p, q;
            rr_pair(m, s, k, p, q);
            if (q >= n) continue;
            const T c = cs[2 * k], sn = cs[2 * k + 1];
            const T x = Es[(size_t)p * n + v], y = Es[(size_t)q * n + v];
            Es[(size_t)p * n + v] = c * x - sn * y;
            Es[(size_t)q * n + v] = sn * x + c * y;
          }
        }
        __syncthreads();
      }
      converged = (sweep_rotated == 0);
    }
    if (my_rot) atomicAdd(rot_total, my_rot);

    // ---- eigenvalues + replay of the reference's selection sort (:207-212, :477-508)
    for (int i = tid; i < n; i += nthr) {
      const T v = Ms[(size_t)i * ldm + i];
      ids[i] = i;
      key[i] = sort_criteria == SORT_DECREASING_EVALS ? v
             : sort_criteria == SORT_INCREASING_EVALS ? -v
             : sort_criteria == SORT_DECREASING_ABS_EVALS ? jabs(v) : -jabs(v);
    }
    __syncthreads();
    const bool do_sort = sort_criteria >= SORT_DECREASING_EVALS && sort_criteria <= SORT_INCREASING_ABS_EVALS;
    if (do_sort && tid < 32) {
      const int lane = tid;
      for (int i = 0; i < n - 1; ++i) {
        // first strict maximum of key[i..n-1]: lowest position wins ties
        T best = key[i]; int pos = i;
        for (int j = i + 1 + lane; j < n; j += 32) {
          const T kj = key[j];
          if (kj > best) { best = kj; pos = j; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          const T ob = __shfl_xor_sync(0xffffffffu, best, off);
          const int op = __shfl_xor_sync(0xffffffffu, pos, off);
          if (ob > best || (ob == best && op < pos)) { best = ob; pos = op; }
        }
        if (lane == 0 && pos != i) {
          const T tk = key[i]; key[i] = key[pos]; key[pos] = tk;
          const int ti = ids[i]; ids[i] = ids[pos]; ids[pos] = ti;
        }
        __syncwarp();
      }
    }
    __syncthreads();

    // ---- store
    if (tid == 0 && n_iters) n_iters[b] = (converged ? *rot_total + 1 : 0);
    for (int k = tid; k < n; k += nthr) {
      const int src = ids[k];
      const T v = Ms[(size_t)src * ldm + src];
      if (soa) evals[(size_t)k * (size_t)batch + (size_t)b] = v;
      else evals[(size_t)b * n + k] = v;
    }
    if (calc_evecs) {
      for (int idx = tid; idx < n * n; idx += nthr) {
        const int k = idx / n, v = idx - k * n;
        const T val = Es[(size_t)ids[k] * n + v];
        if (soa) evecs[(size_t)idx * (size_t)batch + (size_t)b] = val;
        else evecs[(size_t)b * n * n + idx] = val;
      }
    }
    __syncthreads();  // M/E/ids are reused by the next matrix of this block
  }
}

#endif  // __CUDACC__

}  // namespace jpd
