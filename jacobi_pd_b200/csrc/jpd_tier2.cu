// jpd_tier2.cu -- see jpd_internal.h; part of libjpd.so (no CPU path: every entry ends in a kernel launch).
#include <algorithm>
#include <mutex>

#include "jpd_internal.h"
#include "jpd_warp.cuh"

namespace jpd_host {
using namespace jpd;

namespace {
// ------------------------------------------------------------------ tier 2 launch (warp per matrix)
constexpr int kWarpThreads = JPD_WARP_THREADS;

template <typename T, int NP>
int launch_warp(const T* mat, T* evals, T* evecs, int* iters, int n, long long batch, int layout,
                int sort, int calc, int sweeps, cudaStream_t st) {
  using Sh = WarpShape<NP>;
  constexpr int warps = kWarpThreads / 32;
  const long long per_block = (long long)warps * Sh::kGroups;
  const long long blocks = (batch + per_block - 1) / per_block;
  if (blocks > 0x7fffffffLL) return JPD_ERR_UNSUPPORTED;
  const size_t smem = Sh::block_bytes(sizeof(T), warps);
  const bool full = (n == NP);
  auto kernel = calc ? (full ? jpd_warp_kernel<T, NP, true, true> : jpd_warp_kernel<T, NP, true, false>)
                     : (full ? jpd_warp_kernel<T, NP, false, true> : jpd_warp_kernel<T, NP, false, false>);
  if (smem > 48 * 1024)
    JPD_CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kernel<<<dim3((unsigned)blocks), dim3(kWarpThreads), smem, st>>>(
      mat, evals, evecs, iters, n, batch, layout == JPD_LAYOUT_SOA ? 1 : 0, sort, sweeps);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  JPD_CK(cudaGetLastError());
  return JPD_OK;
}

}  // namespace

template <typename T>
int launch_tier2(JPD_TIER_ARGS(T)) {
  if (n <= 8) return launch_warp<T, 8>(mat, evals, evecs, iters, n, batch, layout, sort, calc, sweeps, st);
  if (n <= 16) return launch_warp<T, 16>(mat, evals, evecs, iters, n, batch, layout, sort, calc, sweeps, st);
  return launch_warp<T, 32>(mat, evals, evecs, iters, n, batch, layout, sort, calc, sweeps, st);
}

template int launch_tier2<double>(JPD_TIER_ARGS(double));
template int launch_tier2<float>(JPD_TIER_ARGS(float));

}  // namespace jpd_host
