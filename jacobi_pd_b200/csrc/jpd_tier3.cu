// jpd_tier3.cu -- see jpd_internal.h; part of libjpd.so (no CPU path: every entry ends in a kernel launch).
#include <algorithm>
#include <mutex>

#include "jpd_internal.h"
#include "jpd_cta.cuh"

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

}  // namespace

template <typename T>
int launch_tier3(JPD_TIER_ARGS(T)) {
  if (n <= 64) return launch_cta<T, 64>(mat, evals, evecs, iters, n, batch, layout, sort, calc, sweeps, st);
  if (n <= 96) return launch_cta<T, 96>(mat, evals, evecs, iters, n, batch, layout, sort, calc, sweeps, st);
  return launch_cta<T, 128>(mat, evals, evecs, iters, n, batch, layout, sort, calc, sweeps, st);
}

template int launch_tier3<double>(JPD_TIER_ARGS(double));
template int launch_tier3<float>(JPD_TIER_ARGS(float));

}  // namespace jpd_host
