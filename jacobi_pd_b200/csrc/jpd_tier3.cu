// jpd_tier3.cu -- see jpd_internal.h; part of libjpd.so (no CPU path: every entry ends in a kernel launch).
#include <algorithm>
#include <mutex>

#include "jpd_internal.h"
#include "jpd_cta.cuh"
#include "jpd_cta4.cuh"
#include <cstdlib>

namespace jpd_host {
using namespace jpd;

namespace {
// ------------------------------------------------------------------ tier 3 launch (CTA per matrix, E in registers)

template <typename T, int NPB>
int launch_cta(const T* mat, T* evals, T* evecs, int* iters, int n, long long batch, int layout,
               int sort, int calc, int sweeps, cudaStream_t st) {
  using Sh = CtaShape<NPB>;
  int dev = 0;
  JPD_CK(cudaGetDevice(&dev));
  DeviceInfo info;
  int rc = device_info(dev, &info);
  if (rc != JPD_OK) return rc;
  const size_t smem = Sh::block_bytes(sizeof(T));
  auto kernel = calc ? jpd_cta_kernel<T, NPB, true> : jpd_cta_kernel<T, NPB, false>;
  if (smem > 48 * 1024)
    JPD_CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  JPD_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, Sh::kThreads, smem));
  if (per_sm < 1) return JPD_ERR_UNSUPPORTED;
  const long long grid = std::min<long long>(batch, (long long)per_sm * info.sm_count);
  kernel<<<dim3((unsigned)grid), dim3(Sh::kThreads), smem, st>>>(
      mat, evals, evecs, iters, n, batch, layout == JPD_LAYOUT_SOA ? 1 : 0, sort, sweeps);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  JPD_CK(cudaGetLastError());
  return JPD_OK;
}

// block odd-even ordering, 4x4 tiles of M in registers for two steps, scaled eigenvector rotations (jpd_cta4.cuh)
template <typename T, int NPB>
int launch_cta4(const T* mat, T* evals, T* evecs, int* iters, int n, long long batch, int layout,
                int sort, int calc, int sweeps, cudaStream_t st) {
  using Sh = Cta4Shape<NPB>;
  int dev = 0;
  JPD_CK(cudaGetDevice(&dev));
  DeviceInfo info;
  int rc = device_info(dev, &info);
  if (rc != JPD_OK) return rc;
  const size_t smem = Sh::block_bytes(sizeof(T));
  auto kernel = calc ? jpd_cta4_kernel<T, NPB, true> : jpd_cta4_kernel<T, NPB, false>;
  if (smem > 48 * 1024)
    JPD_CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  JPD_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, Sh::kThreads, smem));
  if (per_sm < 1) return JPD_ERR_UNSUPPORTED;
  const long long grid = std::min<long long>(batch, (long long)per_sm * info.sm_count);
  kernel<<<dim3((unsigned)grid), dim3(Sh::kThreads), smem, st>>>(
      mat, evals, evecs, iters, n, batch, layout == JPD_LAYOUT_SOA ? 1 : 0, sort, sweeps);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  JPD_CK(cudaGetLastError());
  return JPD_OK;
}

// development switch (A/B timing against the first-generation kernel): JPD_TIER3_LEGACY=1
bool tier3_legacy() {
  static const bool v = [] { const char* e = std::getenv("JPD_TIER3_LEGACY"); return e && e[0] == '1'; }();
  return v;
}
bool tier3_force_block() {   // JPD_TIER3_LEGACY=0: the block-ordering kernel at every size (tuning of the switch points)
  static const bool v = [] { const char* e = std::getenv("JPD_TIER3_LEGACY"); return e && e[0] == '0'; }();
  return v;
}

}  // namespace

template <typename T>
int launch_tier3(JPD_TIER_ARGS(T)) {
  // Measured on B200 (tests/scratch/t3sweep.py, t3sweep2.py; DESIGN.md "tier 3"): the block-ordering kernel wins
  // for fp32 at every size (1.22x-1.71x) and for fp64 at 47 <= n <= 64 (1.02x-1.13x) and n >= 97 (1.17x-1.22x);
  // for the other fp64 sizes the first-generation kernel keeps two (65..96) or four resident blocks per SM where
  // the new one, with its 4x4 register tiles, fits fewer, and stays ahead by 3-17 %.
  const bool block_order = !tier3_legacy() && (tier3_force_block() || sizeof(T) == 4 || n >= 97 || (n >= 47 && n <= 64));
  if (block_order) {
    if (n <= 64) return launch_cta4<T, 64>(mat, evals, evecs, iters, n, batch, layout, sort, calc, sweeps, st);
    if constexpr (sizeof(T) == 4) {
      if (n <= 96) return launch_cta4<T, 96>(mat, evals, evecs, iters, n, batch, layout, sort, calc, sweeps, st);
    }
    return launch_cta4<T, 128>(mat, evals, evecs, iters, n, batch, layout, sort, calc, sweeps, st);
  }
  if (n <= 64) return launch_cta<T, 64>(mat, evals, evecs, iters, n, batch, layout, sort, calc, sweeps, st);
  if (n <= 96) return launch_cta<T, 96>(mat, evals, evecs, iters, n, batch, layout, sort, calc, sweeps, st);
  return launch_cta<T, 128>(mat, evals, evecs, iters, n, batch, layout, sort, calc, sweeps, st);
}

template int launch_tier3<double>(JPD_TIER_ARGS(double));
template int launch_tier3<float>(JPD_TIER_ARGS(float));

}  // namespace jpd_host
