// jpd_tier1.cu -- see jpd_internal.h; part of libjpd.so (no CPU path: every entry ends in a kernel launch).
#include <algorithm>
#include <mutex>

#include "jpd_internal.h"
#include "jpd_small.cuh"
#include "jpd_fused.cuh"

namespace jpd_host {
using namespace jpd;

namespace {
constexpr int kSmallThreads = JPD_SMALL_THREADS;

template <int N, typename T, bool SOA, bool EVECS>
int launch_small_layout(const T* mat, T* evals, T* evecs, int* iters, long long batch, int sort, int sweeps,
                        cudaStream_t st) {
  const long long tiles = (batch + kSmallThreads - 1) / kSmallThreads;
  if (tiles > 0x7fffffffLL) return JPD_ERR_UNSUPPORTED;
  jpd_small_kernel<N, T, SOA, EVECS><<<dim3((unsigned)tiles), dim3(kSmallThreads), 0, st>>>(mat, evals, evecs, iters,
                                                                                           batch, sort, sweeps);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  JPD_CK(cudaGetLastError());
  return JPD_OK;
}

template <int N, typename T>
int launch_small(const T* mat, T* evals, T* evecs, int* iters, long long batch, int layout,
                 int sort, int calc, int sweeps, cudaStream_t st) {
  if (layout == JPD_LAYOUT_SOA) {
    return calc ? launch_small_layout<N, T, true, true>(mat, evals, evecs, iters, batch, sort, sweeps, st)
                : launch_small_layout<N, T, true, false>(mat, evals, evecs, iters, batch, sort, sweeps, st);
  }
  return calc ? launch_small_layout<N, T, false, true>(mat, evals, evecs, iters, batch, sort, sweeps, st)
              : launch_small_layout<N, T, false, false>(mat, evals, evecs, iters, batch, sort, sweeps, st);
}

}  // namespace

template <typename T>
int launch_tier1(JPD_TIER_ARGS(T)) {
  switch (n) {
#ifndef JPD_VARIANT_BUILD
    case 1: return launch_small<1, T>(mat, evals, evecs, iters, batch, layout, sort, calc, sweeps, st);
    case 2: return launch_small<2, T>(mat, evals, evecs, iters, batch, layout, sort, calc, sweeps, st);
#endif
    case 3: return launch_small<3, T>(mat, evals, evecs, iters, batch, layout, sort, calc, sweeps, st);
    case 4: return launch_small<4, T>(mat, evals, evecs, iters, batch, layout, sort, calc, sweeps, st);
#ifndef JPD_VARIANT_BUILD
    case 5: return launch_small<5, T>(mat, evals, evecs, iters, batch, layout, sort, calc, sweeps, st);
    case 6: return launch_small<6, T>(mat, evals, evecs, iters, batch, layout, sort, calc, sweeps, st);
#endif
    default: return JPD_ERR_UNSUPPORTED;
  }
  static_assert(kSmallMaxN == 6, "switch above lists the tier-1 sizes");
}

// ------------------------------------------------------------------ fused producers
template <typename T>
int quaternion_device(const T* corr, T* lambda, T* quat, int* iters, long long batch, int layout,
                      int all, int sweeps, cudaStream_t st) {
  if (batch < 0 || (layout != JPD_LAYOUT_AOS && layout != JPD_LAYOUT_SOA)) return JPD_ERR_INVALID_ARG;
  if (batch == 0) return JPD_OK;
  if (!corr || !lambda || !quat) return JPD_ERR_INVALID_ARG;
  if (sweeps < 0) sweeps = 0;
  const long long blocks = (batch + kSmallThreads - 1) / kSmallThreads;
  if (blocks > 0x7fffffffLL) return JPD_ERR_UNSUPPORTED;
  const dim3 grid((unsigned)blocks), block(kSmallThreads);
  const bool soa = layout == JPD_LAYOUT_SOA;
  if (all) {
    if (soa) jpd_quat_kernel<T, true, true><<<grid, block, 0, st>>>(corr, lambda, quat, iters, batch, sweeps);
    else     jpd_quat_kernel<T, false, true><<<grid, block, 0, st>>>(corr, lambda, quat, iters, batch, sweeps);
  } else {
    if (soa) jpd_quat_kernel<T, true, false><<<grid, block, 0, st>>>(corr, lambda, quat, iters, batch, sweeps);
    else     jpd_quat_kernel<T, false, false><<<grid, block, 0, st>>>(corr, lambda, quat, iters, batch, sweeps);
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  JPD_CK(cudaGetLastError());
  return JPD_OK;
}

template <typename T>
int principal_axes_device(const T* inertia6, T* moments, T* axes, int* iters, long long batch,
                          int layout, int sweeps, cudaStream_t st) {
  if (batch < 0 || (layout != JPD_LAYOUT_AOS && layout != JPD_LAYOUT_SOA)) return JPD_ERR_INVALID_ARG;
  if (batch == 0) return JPD_OK;
  if (!inertia6 || !moments || !axes) return JPD_ERR_INVALID_ARG;
  if (sweeps < 0) sweeps = 0;
  const long long blocks = (batch + kSmallThreads - 1) / kSmallThreads;
  if (blocks > 0x7fffffffLL) return JPD_ERR_UNSUPPORTED;
  const dim3 grid((unsigned)blocks), block(kSmallThreads);
  if (layout == JPD_LAYOUT_SOA) jpd_inertia_kernel<T, true><<<grid, block, 0, st>>>(inertia6, moments, axes, iters, batch, sweeps);
  else                          jpd_inertia_kernel<T, false><<<grid, block, 0, st>>>(inertia6, moments, axes, iters, batch, sweeps);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  JPD_CK(cudaGetLastError());
  return JPD_OK;
}


template int launch_tier1<double>(JPD_TIER_ARGS(double));
template int launch_tier1<float>(JPD_TIER_ARGS(float));
template int quaternion_device<double>(const double*, double*, double*, int*, long long, int, int, int, cudaStream_t);
template int quaternion_device<float>(const float*, float*, float*, int*, long long, int, int, int, cudaStream_t);
template int principal_axes_device<double>(const double*, double*, double*, int*, long long, int, int, cudaStream_t);
template int principal_axes_device<float>(const float*, float*, float*, int*, long long, int, int, cudaStream_t);

}  // namespace jpd_host
