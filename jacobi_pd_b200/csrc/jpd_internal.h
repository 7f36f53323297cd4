// jpd_internal.h -- what the translation units of libjpd.so share: error plumbing, the
// per-device facts, and one launch entry per kernel tier (each tier is compiled in its own
// .cu so that the library builds in parallel; explicit instantiations for float and double).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <string>

#include "jpd.h"

namespace jpd_host {

constexpr int kMaxDevices = 64;
constexpr int kSmallMaxN = 6;      // tier 1 upper bound (registers only)
constexpr int kWarpMaxN = 32;      // tier 2
constexpr int kCtaMaxN = 128;      // tier 3
constexpr int kClusterMaxN = 256;  // tier 4

void set_last_cuda_error(const char* expr, cudaError_t e);
extern std::atomic<long long> g_launches;

#define JPD_CK(expr)                                                \
  do {                                                              \
    cudaError_t _e = (expr);                                        \
    if (_e != cudaSuccess) {                                        \
      ::jpd_host::set_last_cuda_error(#expr, _e);                   \
      return JPD_ERR_CUDA;                                          \
    }                                                               \
  } while (0)

struct DeviceInfo {
  bool ready = false;
  int sm_count = 148;
  int smem_optin = 232448;
  bool pool_tuned = false;
};
int device_info(int dev, DeviceInfo* out);
void keep_freed_workspace_cached(int dev);   // stream-ordered pool: release threshold = never

// mat/evals/evecs/iters: device pointers; layout JPD_LAYOUT_*; arguments already validated.
#define JPD_TIER_ARGS(T) const T* mat, T* evals, T* evecs, int* iters, int n, long long batch, int layout, \
                         int sort, int calc, int sweeps, cudaStream_t st
template <typename T> int launch_tier1(JPD_TIER_ARGS(T));   // n <= 6: thread per matrix
template <typename T> int launch_tier2(JPD_TIER_ARGS(T));   // n <= 32: warp per matrix
template <typename T> int launch_tier3(JPD_TIER_ARGS(T));   // n <= 128: CTA per matrix
template <typename T> int launch_tier4(JPD_TIER_ARGS(T));   // n <= 256: 4-CTA cluster per matrix
template <typename T> int launch_tier5(JPD_TIER_ARGS(T));   // n <= JPD_MAX_N
template <typename T> int quaternion_device(const T* corr, T* lambda, T* quat, int* iters, long long batch, int layout,
                                            int all, int sweeps, cudaStream_t st);
template <typename T> int principal_axes_device(const T* inertia6, T* moments, T* axes, int* iters, long long batch,
                                                int layout, int sweeps, cudaStream_t st);
template <typename T> int convert_layout(const void* src, int src_layout, T* dst, int dst_layout, int n, long long batch,
                                         cudaStream_t st);

}  // namespace jpd_host
