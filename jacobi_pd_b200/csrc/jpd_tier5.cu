// jpd_tier5.cu -- see jpd_internal.h; part of libjpd.so (no CPU path: every entry ends in a kernel launch).
#include <algorithm>
#include <mutex>

#include "jpd_internal.h"
#include "jpd_block.cuh"
#include "jpd_layout.cuh"

namespace jpd_host {
using namespace jpd;

namespace {
// ------------------------------------------------------------------ tier 5 plan + launch
BlockPlan make_block_plan(int n, int elem, int calc, int smem_limit) {
  BlockPlan p{};
  const int m = n + (n & 1), half = m / 2;
  p.ldm = n | 1;
  const size_t base = (size_t)(2 * half + n) * elem + (size_t)(n + 4) * sizeof(int) + 16;
  const size_t mb = (size_t)n * p.ldm * elem;
  const size_t eb = calc ? (size_t)n * n * elem : 0;
  if (base + mb + eb <= (size_t)smem_limit) { p.m_in_smem = 1; p.e_in_smem = 1; }
  else if (base + mb <= (size_t)smem_limit) { p.m_in_smem = 1; p.e_in_smem = 0; }
  else { p.m_in_smem = 0; p.e_in_smem = 0; }
  p.smem_bytes = base + (p.m_in_smem ? mb : 0) + (p.e_in_smem ? eb : 0);
  p.work_elems = (p.m_in_smem ? 0 : (size_t)n * p.ldm) + ((p.e_in_smem || !calc) ? 0 : (size_t)n * n);
  const long long items = (long long)half * (half - 1) / 2 + (calc ? (long long)half * n : 0);
  long long thr = ((items / 4) + 31) / 32 * 32;
  p.threads = (int)std::min<long long>(1024, std::max<long long>(32, thr));
  return p;
}

template <typename T>
int launch_block(const T* mat, T* evals, T* evecs, int* iters, int n, long long batch, int layout,
                 int sort, int calc, int sweeps, cudaStream_t st) {
  int dev = 0;
  JPD_CK(cudaGetDevice(&dev));
  DeviceInfo info;
  int rc = device_info(dev, &info);
  if (rc != JPD_OK) return rc;
  const BlockPlan plan = make_block_plan(n, (int)sizeof(T), calc, info.smem_optin);
  auto kernel = jpd_block_kernel<T>;
  if (plan.smem_bytes > 48 * 1024)
    JPD_CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem_bytes));
  int per_sm = 1;
  JPD_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, plan.threads, plan.smem_bytes));
  if (per_sm < 1) return JPD_ERR_UNSUPPORTED;
  const long long resident = (long long)per_sm * info.sm_count;
  const long long grid = std::min<long long>(batch, resident);

  T* work = nullptr;
  if (plan.work_elems) {
    keep_freed_workspace_cached(dev);
    JPD_CK(cudaMallocAsync((void**)&work, plan.work_elems * sizeof(T) * (size_t)grid, st));
  }
  kernel<<<dim3((unsigned)grid), dim3(plan.threads), plan.smem_bytes, st>>>(
      mat, evals, evecs, iters, n, batch, layout == JPD_LAYOUT_SOA ? 1 : 0, sort, calc, sweeps,
      plan.ldm, plan.m_in_smem, plan.e_in_smem, work, plan.work_elems);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t le = cudaGetLastError();
  if (work) cudaFreeAsync(work, st);
  JPD_CK(le);
  return JPD_OK;
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
