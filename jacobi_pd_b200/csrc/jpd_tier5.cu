// jpd_tier5.cu -- see jpd_internal.h; part of libjpd.so (no CPU path: every entry ends in a kernel launch).
#include <algorithm>
#include <mutex>

#include "jpd_internal.h"
#include <cstdlib>
#include <cstring>
#include <vector>

#include "jpd_blockjacobi.cuh"
#include "jpd_layout.cuh"

namespace jpd_host {
using namespace jpd;

namespace {
// ------------------------------------------------------------------ tier 5: block Jacobi (jpd_blockjacobi.cuh)
// round-robin tournament of nb (even) block indices: step s in [0, nb-1) -> nb/2 disjoint pairs
void block_pairs(int nb, int step, std::vector<int2>& out) {
  out.clear();
  const int m = nb - 1;
  for (int k = 0; k < nb / 2; ++k) {
    int a = (k == 0) ? m : (step + k) % m;
    int b = (k == 0) ? step : (step - k + m) % m;
    if (a > b) std::swap(a, b);
    out.push_back(make_int2(a, b));
  }
}

// Inner sweeps of a pivot solve.  Full convergence of every 64 x 64 block is not needed for the outer
// iteration to converge quadratically: measured on B200 (tests/scratch/t5inner.py), a cap of 2 inner
// sweeps gives the same accuracy and outer sweep count + 1 at 1.25x-1.28x the speed of full inner
// convergence (n = 300: 1436 -> 1845 matrices/s, 512: 435 -> 544, 1000: 53.9 -> 68.3).  The cap is
// passed as a NEGATIVE sweep budget, which makes an unfinished solve report -(rotations + 1) instead of
// 0, so the rotation count stays a count.  JPD_BLOCK_INNER_SWEEPS overrides (tuning aid).
int inner_sweeps() {
  static int v = -1;
  if (v < 0) {
    v = 2;
    if (const char* env = std::getenv("JPD_BLOCK_INNER_SWEEPS")) {
      const long x = std::strtol(env, nullptr, 10);
      if (x >= 1 && x <= 50) v = (int)x;
    }
  }
  return v;
}

template <typename T>
int launch_block(const T* mat, T* evals, T* evecs, int* iters, int n, long long batch, int layout,
                 int sort, int calc, int sweeps, cudaStream_t st) {
  int dev = 0;
  JPD_CK(cudaGetDevice(&dev));
  keep_freed_workspace_cached(dev);
  const int np = (n + kBjPivot - 1) / kBjPivot * kBjPivot;
  const int nb = np / kBjBlock, npairs = nb / 2, nsteps = nb - 1;
  const size_t mat_elems = (size_t)np * np;
  const int soa = layout == JPD_LAYOUT_SOA ? 1 : 0;

  // matrices per pass: workspace (M, E, pivots, Q) capped at ~6 GB
  const size_t per_matrix = sizeof(T) * (2 * mat_elems + (size_t)npairs * (2 * 4096 + 64)) + 64;
  long long group = std::max<long long>(1, (long long)(((size_t)6 << 30) / per_matrix));
  group = std::min(group, batch);

  T *Mw = nullptr, *Ew = nullptr, *P = nullptr, *Q = nullptr, *pev = nullptr;
  int *piv_it = nullptr, *counters = nullptr;
  int2* d_pairs = nullptr;
  int2* h_pairs = nullptr;
  int* h_flag = nullptr;
  int status = JPD_OK;
  cudaError_t err = cudaSuccess;
  const char* what = "";
#define BJ_TRY(expr) do { if (err == cudaSuccess) { err = (expr); if (err != cudaSuccess) what = #expr; } } while (0)
  BJ_TRY(cudaMallocAsync((void**)&Mw, sizeof(T) * mat_elems * group, st));
  BJ_TRY(cudaMallocAsync((void**)&Ew, sizeof(T) * mat_elems * group, st));
  BJ_TRY(cudaMallocAsync((void**)&P, sizeof(T) * 4096 * npairs * group, st));
  BJ_TRY(cudaMallocAsync((void**)&Q, sizeof(T) * 4096 * npairs * group, st));
  BJ_TRY(cudaMallocAsync((void**)&pev, sizeof(T) * 64 * npairs * group, st));
  BJ_TRY(cudaMallocAsync((void**)&piv_it, sizeof(int) * npairs * group, st));
  BJ_TRY(cudaMallocAsync((void**)&counters, sizeof(int) * (3 * group + 1), st));   // rot_total | rot_sweep | converged | any_left
  BJ_TRY(cudaMallocAsync((void**)&d_pairs, sizeof(int2) * npairs * nsteps, st));
  BJ_TRY(cudaMallocHost((void**)&h_pairs, sizeof(int2) * npairs * nsteps));
  BJ_TRY(cudaMallocHost((void**)&h_flag, sizeof(int)));
  if (err == cudaSuccess) {
    std::vector<int2> pr;
    for (int s = 0; s < nsteps; ++s) {
      block_pairs(nb, s, pr);
      std::memcpy(h_pairs + (size_t)s * npairs, pr.data(), sizeof(int2) * npairs);
    }
    BJ_TRY(cudaMemcpyAsync(d_pairs, h_pairs, sizeof(int2) * npairs * nsteps, cudaMemcpyHostToDevice, st));
  }
  const size_t ortho_smem = sizeof(T) * 3 * kBjPivot * kBjLd;
  const size_t apply_smem = sizeof(T) * JPD_BJ_APPLY_BUFFERS * kBjPivot * kBjLd;
  auto apply = calc ? blk_apply_kernel<T, true> : blk_apply_kernel<T, true>;   // E is always kept: it marks the padding rows
  BJ_TRY(cudaFuncSetAttribute(apply, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)apply_smem));
  BJ_TRY(cudaFuncSetAttribute(blk_orthonormalize_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ortho_smem));
  const int ntile = npairs * (npairs + 1) / 2;
  const int items = ntile + npairs * (np / kBjPivot);

  for (long long b0 = 0; b0 < batch && err == cudaSuccess && status == JPD_OK; b0 += group) {
    const long long cnt = std::min(group, batch - b0);
    int* rot_total = counters; int* rot_sweep = counters + group; int* converged = counters + 2 * group;
    int* any_left = counters + 3 * group;
    BJ_TRY(cudaMemsetAsync(counters, 0, sizeof(int) * (3 * group + 1), st));
    const long long init_blocks = std::min<long long>((cnt * (long long)mat_elems + 255) / 256, 148 * 64);
    blk_init_kernel<T><<<(unsigned)init_blocks, 256, 0, st>>>(mat, Mw, Ew, n, np, batch, b0, cnt, soa);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    const unsigned cb = (unsigned)((cnt + 127) / 128);
    for (int sweep = 0; sweep < sweeps && err == cudaSuccess && status == JPD_OK; ++sweep) {
      BJ_TRY(cudaMemsetAsync(any_left, 0, sizeof(int), st));
      for (int s = 0; s < nsteps && status == JPD_OK; ++s) {
        const int2* prs = d_pairs + (size_t)s * npairs;
        blk_gather_kernel<T><<<dim3(npairs, (unsigned)cnt), kBjThreads, 0, st>>>(Mw, P, prs, npairs, np);
        // SORTED pivot solves: the rows of Q must come in a consistent order (here: decreasing
        // eigenvalue).  With an arbitrary order -- the odd-even ordering of the inner solver
        // reverses the positions every sweep -- Q carries a large permutation, the diagonal
        // never settles and block Jacobi converges only linearly (measured: n = 1000 not done
        // after 50 sweeps; sorted: 8 sweeps).  Sorted, Q tends to the identity.
        status = launch_tier3<T>(P, pev, Q, piv_it, kBjPivot, (long long)npairs * cnt, JPD_LAYOUT_AOS,
                                 JPD_SORT_DECREASING_EVALS, 1, -inner_sweeps(), st);
        if (status != JPD_OK) break;
        blk_orthonormalize_kernel<T><<<(unsigned)(npairs * cnt), kBjGemmThreads, ortho_smem, st>>>(Q);
        apply<<<dim3(items, (unsigned)cnt), kBjGemmThreads, apply_smem, st>>>(Mw, Ew, Q, prs, npairs, np);
        blk_count_kernel<<<cb, 128, 0, st>>>(piv_it, npairs, cnt, rot_total, rot_sweep, any_left);
        g_launches.fetch_add(4, std::memory_order_relaxed);
      }
      blk_sweep_end_kernel<<<cb, 128, 0, st>>>(rot_sweep, converged, cnt);
      g_launches.fetch_add(1, std::memory_order_relaxed);
      BJ_TRY(cudaGetLastError());
      BJ_TRY(cudaMemcpyAsync(h_flag, any_left, sizeof(int), cudaMemcpyDeviceToHost, st));
      BJ_TRY(cudaStreamSynchronize(st));     // one flag per outer sweep
      if (err == cudaSuccess && *h_flag == 0) break;   // a whole sweep without a rotation, for every matrix
    }
    if (err != cudaSuccess || status != JPD_OK) break;
    const size_t fin_smem = sizeof(T) * np + sizeof(int) * 2 * np;
    if (calc) blk_finalize_kernel<T, true><<<(unsigned)cnt, 256, fin_smem, st>>>(Mw, Ew, rot_total, converged, evals, evecs, iters,
                                                                                n, np, batch, b0, soa, sort, sweeps);
    else      blk_finalize_kernel<T, false><<<(unsigned)cnt, 256, fin_smem, st>>>(Mw, Ew, rot_total, converged, evals, evecs, iters,
                                                                                 n, np, batch, b0, soa, sort, sweeps);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    BJ_TRY(cudaGetLastError());
  }
#undef BJ_TRY
  if (h_pairs || h_flag) cudaStreamSynchronize(st);   // the pinned scratch must outlive the copies that use it
  if (Mw) cudaFreeAsync(Mw, st);
  if (Ew) cudaFreeAsync(Ew, st);
  if (P) cudaFreeAsync(P, st);
  if (Q) cudaFreeAsync(Q, st);
  if (pev) cudaFreeAsync(pev, st);
  if (piv_it) cudaFreeAsync(piv_it, st);
  if (counters) cudaFreeAsync(counters, st);
  if (d_pairs) cudaFreeAsync(d_pairs, st);
  if (h_pairs) cudaFreeHost(h_pairs);
  if (h_flag) cudaFreeHost(h_flag);
  if (err != cudaSuccess) { set_last_cuda_error(what, err); return JPD_ERR_CUDA; }
  return status;
}

}  // namespace

template <typename T>
int launch_tier5(JPD_TIER_ARGS(T)) {
  return launch_block<T>(mat, evals, evecs, iters, n, batch, layout, sort, calc, sweeps, st);
}

// ------------------------------------------------------------------ layout adapters
template <typename T>
int convert_layout(const void* src, int src_layout, T* dst, int dst_layout, int n, long long batch,
                   cudaStream_t st) {
  if (n < 1 || n > JPD_MAX_N || batch < 0) return JPD_ERR_INVALID_ARG;
  if (src_layout < kLayoutAoS || src_layout > kLayoutPointers) return JPD_ERR_INVALID_ARG;
  if (dst_layout < kLayoutAoS || dst_layout > kLayoutPackedSoA) return JPD_ERR_INVALID_ARG;
  if (batch == 0) return JPD_OK;
  if (!src || !dst) return JPD_ERR_INVALID_ARG;
  const int nn = n * n;
  if (nn <= 256 && ((src_layout == kLayoutAoS && dst_layout == kLayoutSoA) || (src_layout == kLayoutSoA && dst_layout == kLayoutAoS))) {
    const int tb_shift = nn <= 32 ? 7 : (nn <= 64 ? 6 : 5), tb = 1 << tb_shift;
    const long long blocks = (batch + tb - 1) / tb;
    if (blocks > 0x7fffffffLL) return JPD_ERR_UNSUPPORTED;
    const size_t smem = sizeof(T) * (size_t)tb * (nn | 1);
    if (src_layout == kLayoutAoS) {
      if (smem > 48 * 1024) JPD_CK(cudaFuncSetAttribute(jpd_convert_flat_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      jpd_convert_flat_kernel<T, true><<<(unsigned)blocks, 256, smem, st>>>(static_cast<const T*>(src), dst, nn, batch, tb, tb_shift);
    } else {
      if (smem > 48 * 1024) JPD_CK(cudaFuncSetAttribute(jpd_convert_flat_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      jpd_convert_flat_kernel<T, false><<<(unsigned)blocks, 256, smem, st>>>(static_cast<const T*>(src), dst, nn, batch, tb, tb_shift);
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    JPD_CK(cudaGetLastError());
    return JPD_OK;
  }
  const long long bx = (batch + 31) / 32;
  const long long by = ((long long)n * n + 31) / 32;
  if (bx > 0x7fffffffLL || by > 65535) return JPD_ERR_UNSUPPORTED;
  jpd_convert_kernel<T><<<dim3((unsigned)bx, (unsigned)by), dim3(32, 8), 0, st>>>(src, src_layout, dst, dst_layout, n, batch);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  JPD_CK(cudaGetLastError());
  return JPD_OK;
}


template int launch_tier5<double>(JPD_TIER_ARGS(double));
template int launch_tier5<float>(JPD_TIER_ARGS(float));
template int convert_layout<double>(const void*, int, double*, int, int, long long, cudaStream_t);
template int convert_layout<float>(const void*, int, float*, int, int, long long, cudaStream_t);

}  // namespace jpd_host
