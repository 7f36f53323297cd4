// jpd_api.cu -- the C ABI of libjpd.so (declared in include/jpd.h): argument checks,
// tier dispatch, the host-buffer streaming pipeline and the multi-GPU slicer.
//
// Replaces, for whole batches, the call path
//   Jacobi<...>::Diagonalize  /root/reference/include/jacobi_pd.hpp:159-220
// There is no CPU fallback anywhere in this file: every path ends in a kernel launch
// or in an error code.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "jpd.h"
#include "jpd_block.cuh"
#include "jpd_small.cuh"
#include "jpd_warp.cuh"
#include "jpd_cta.cuh"
#include "jpd_cluster.cuh"
#include "jpd_fused.cuh"
#include "jpd_layout.cuh"

namespace {

using namespace jpd;

constexpr int kSmallMaxN = 6;      // tier 1 upper bound (registers only)
constexpr int kSmallThreads = JPD_SMALL_THREADS;
constexpr int kMaxDevices = 64;

thread_local std::string g_last_cuda_error;
std::atomic<long long> g_launches{0};

#define JPD_CK(expr)                                                                   \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      g_last_cuda_error = std::string(#expr) + ": " + cudaGetErrorString(_e);          \
      return JPD_ERR_CUDA;                                                             \
    }                                                                                  \
  } while (0)

struct DeviceInfo {
  bool ready = false;
  int sm_count = 148;
  int smem_optin = 232448;
  bool pool_tuned = false;
};
DeviceInfo g_dev[kMaxDevices];
std::mutex g_dev_mutex;

int device_info(int dev, DeviceInfo* out) {
  if (dev < 0 || dev >= kMaxDevices) return JPD_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lock(g_dev_mutex);
  DeviceInfo& d = g_dev[dev];
  if (!d.ready) {
    JPD_CK(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev));
    JPD_CK(cudaDeviceGetAttribute(&d.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    d.ready = true;
  }
  *out = d;
  return JPD_OK;
}

// ------------------------------------------------------------------ tier 1 launch
template <int N, typename T>
int launch_small(const T* mat, T* evals, T* evecs, int* iters, long long batch, int layout,
                 int sort, int calc, int sweeps, cudaStream_t st) {
  const long long blocks = (batch + kSmallThreads - 1) / kSmallThreads;
  if (blocks > 0x7fffffffLL) return JPD_ERR_UNSUPPORTED;
  const dim3 grid((unsigned)blocks), block(kSmallThreads);
  if (layout == JPD_LAYOUT_SOA) {
    if (calc) jpd_small_kernel<N, T, true, true><<<grid, block, 0, st>>>(mat, evals, evecs, iters, batch, sort, sweeps);
    else      jpd_small_kernel<N, T, true, false><<<grid, block, 0, st>>>(mat, evals, evecs, iters, batch, sort, sweeps);
  } else {
    if (calc) jpd_small_kernel<N, T, false, true><<<grid, block, 0, st>>>(mat, evals, evecs, iters, batch, sort, sweeps);
    else      jpd_small_kernel<N, T, false, false><<<grid, block, 0, st>>>(mat, evals, evecs, iters, batch, sort, sweeps);
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  JPD_CK(cudaGetLastError());
  return JPD_OK;
}

// ------------------------------------------------------------------ tier 2 launch (warp per matrix)
constexpr int kWarpMaxN = 32;
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

// ------------------------------------------------------------------ tier 3 launch (CTA per matrix, E in registers)
constexpr int kCtaMaxN = 128;

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

// ------------------------------------------------------------------ tier 4 launch (4-CTA cluster per matrix)
constexpr int kClusterMaxN = ClusterShape::kN;

template <typename T>
int launch_cluster(const T* mat, T* evals, T* evecs, int* iters, int n, long long batch, int layout,
                   int sort, int calc, int sweeps, cudaStream_t st) {
  using Sh = ClusterShape;
  const size_t smem = Sh::block_bytes(sizeof(T));
  auto kernel = calc ? jpd_cluster_kernel<T, true> : jpd_cluster_kernel<T, false>;
  JPD_CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int max_clusters = 0;
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(Sh::kCtas * 64);
    cfg.blockDim = dim3(Sh::kThreads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = Sh::kCtas; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, kernel, &cfg) != cudaSuccess || max_clusters < 1) {
      (void)cudaGetLastError();
      max_clusters = 32;
    }
  }
  const long long clusters = std::min<long long>(batch, max_clusters);
  kernel<<<dim3((unsigned)(clusters * Sh::kCtas)), dim3(Sh::kThreads), smem, st>>>(
      mat, evals, evecs, iters, n, batch, layout == JPD_LAYOUT_SOA ? 1 : 0, sort, sweeps);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  JPD_CK(cudaGetLastError());
  return JPD_OK;
}

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
    {
      std::lock_guard<std::mutex> lock(g_dev_mutex);
      if (!g_dev[dev].pool_tuned) {  // keep freed workspace cached in the stream-ordered pool
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
          unsigned long long keep = ~0ull;
          cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        g_dev[dev].pool_tuned = true;
      }
    }
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

// ------------------------------------------------------------------ device entry
template <typename T>
int diagonalize_device(const T* mat, T* evals, T* evecs, int* iters, int n, long long batch,
                       int layout, int sort, int calc, int sweeps, cudaStream_t st) {
  if (n < 1 || n > JPD_MAX_N || batch < 0) return JPD_ERR_INVALID_ARG;
  if (layout != JPD_LAYOUT_AOS && layout != JPD_LAYOUT_SOA) return JPD_ERR_INVALID_ARG;
  if (batch == 0) return JPD_OK;
  if (!mat || !evals || (calc && !evecs)) return JPD_ERR_INVALID_ARG;
  if (sweeps < 0) sweeps = 0;
  calc = calc ? 1 : 0;
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
    default: break;
  }
  static_assert(kSmallMaxN == 6, "switch above lists the tier-1 sizes");
  if (n <= 8) return launch_warp<T, 8>(mat, evals, evecs, iters, n, batch, layout, sort, calc, sweeps, st);
  if (n <= 16) return launch_warp<T, 16>(mat, evals, evecs, iters, n, batch, layout, sort, calc, sweeps, st);
  if (n <= kWarpMaxN) return launch_warp<T, 32>(mat, evals, evecs, iters, n, batch, layout, sort, calc, sweeps, st);
  if (n <= 64) return launch_cta<T, 64>(mat, evals, evecs, iters, n, batch, layout, sort, calc, sweeps, st);
  if (n <= 96) return launch_cta<T, 96>(mat, evals, evecs, iters, n, batch, layout, sort, calc, sweeps, st);
  if (n <= kCtaMaxN) return launch_cta<T, 128>(mat, evals, evecs, iters, n, batch, layout, sort, calc, sweeps, st);
  if (n <= kClusterMaxN) return launch_cluster<T>(mat, evals, evecs, iters, n, batch, layout, sort, calc, sweeps, st);
  return launch_block<T>(mat, evals, evecs, iters, n, batch, layout, sort, calc, sweeps, st);
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

// ------------------------------------------------------------------ host pipeline
// Per-device staging: 3 slots x (in, evals, evecs, iters) + 3 streams, grown on demand and
// kept until jpd_release().  One host call at a time per device (mutex).
struct Slot {
  cudaStream_t stream = nullptr;
  void* in = nullptr; void* evals = nullptr; void* evecs = nullptr; int* iters = nullptr;
  size_t in_b = 0, evals_b = 0, evecs_b = 0, iters_b = 0;
};
struct Pipeline {
  std::mutex mutex;
  Slot slot[3];
  // small-call path (batch = 1 through the drop-in class): one pinned + one device buffer
  void* small_pin = nullptr;
  void* small_dev = nullptr;
};
constexpr size_t kSmallCallBytes = (size_t)1 << 20;
Pipeline g_pipe[kMaxDevices];

int grow(void** p, size_t* have, size_t want) {
  if (*have >= want) return JPD_OK;
  if (*p) JPD_CK(cudaFree(*p));
  *p = nullptr; *have = 0;
  JPD_CK(cudaMalloc(p, want));
  *have = want;
  return JPD_OK;
}

// host_pitch: elements between consecutive SoA planes in the HOST arrays (== the caller's
// full batch, also when only a slice [first, first+batch) is processed).
template <typename T>
int diagonalize_host(const T* h_mat, T* h_evals, T* h_evecs, int* h_iters, int n, long long batch,
                     long long host_pitch, int layout, int sort, int calc, int sweeps, int device) {
  if (n < 1 || n > JPD_MAX_N || batch < 0) return JPD_ERR_INVALID_ARG;
  if (layout != JPD_LAYOUT_AOS && layout != JPD_LAYOUT_SOA) return JPD_ERR_INVALID_ARG;
  if (device < 0 || device >= kMaxDevices) return JPD_ERR_INVALID_ARG;
  if (batch == 0) return JPD_OK;
  if (!h_mat || !h_evals || (calc && !h_evecs)) return JPD_ERR_INVALID_ARG;
  JPD_CK(cudaSetDevice(device));
  Pipeline& pipe = g_pipe[device];
  std::lock_guard<std::mutex> lock(pipe.mutex);

  const size_t nn = (size_t)n * n;
  const size_t per_matrix = sizeof(T) * (nn + n + (calc ? nn : 0)) + 4;

  // ---- small calls (the drop-in scalar class: batch = 1): everything through ONE pinned
  // staging buffer -- one H2D copy, the kernel, one D2H copy of evals|evecs|n_iters, one
  // synchronisation -- instead of the chunked three-stream pipeline (5 copies, 3 syncs)
  {
    auto up16 = [](size_t x) { return (x + 15) / 16 * 16; };
    const size_t in_b = sizeof(T) * nn * batch, ev_b = sizeof(T) * n * batch;
    const size_t vec_b = calc ? sizeof(T) * nn * batch : 0, it_b = sizeof(int) * batch;
    const size_t ev_off = up16(in_b), vec_off = ev_off + up16(ev_b), it_off = vec_off + up16(vec_b);
    const size_t total = it_off + it_b;
    if (total <= kSmallCallBytes && host_pitch == batch) {
      if (!pipe.small_pin) {
        JPD_CK(cudaHostAlloc(&pipe.small_pin, kSmallCallBytes, cudaHostAllocDefault));
        JPD_CK(cudaMalloc(&pipe.small_dev, kSmallCallBytes));
      }
      if (!pipe.slot[0].stream) JPD_CK(cudaStreamCreateWithFlags(&pipe.slot[0].stream, cudaStreamNonBlocking));
      cudaStream_t st = pipe.slot[0].stream;
      char* hp = static_cast<char*>(pipe.small_pin);
      char* dp = static_cast<char*>(pipe.small_dev);
      std::memcpy(hp, h_mat, in_b);
      JPD_CK(cudaMemcpyAsync(dp, hp, in_b, cudaMemcpyHostToDevice, st));
      const int rc = diagonalize_device<T>(reinterpret_cast<const T*>(dp), reinterpret_cast<T*>(dp + ev_off),
                                           calc ? reinterpret_cast<T*>(dp + vec_off) : nullptr,
                                           reinterpret_cast<int*>(dp + it_off), n, batch, layout, sort, calc, sweeps, st);
      if (rc != JPD_OK) return rc;
      JPD_CK(cudaMemcpyAsync(hp + ev_off, dp + ev_off, total - ev_off, cudaMemcpyDeviceToHost, st));
      JPD_CK(cudaStreamSynchronize(st));
      std::memcpy(h_evals, hp + ev_off, ev_b);
      if (calc) std::memcpy(h_evecs, hp + vec_off, vec_b);
      if (h_iters) std::memcpy(h_iters, hp + it_off, it_b);
      return JPD_OK;
    }
  }
  // bytes per pipeline chunk (in + out): 64 MB by default, JPD_HOST_CHUNK_MB overrides (tuning aid)
  size_t chunk_bytes = (size_t)64 << 20;
  if (const char* env = std::getenv("JPD_HOST_CHUNK_MB")) {
    const long mb = std::strtol(env, nullptr, 10);
    if (mb >= 1 && mb <= 4096) chunk_bytes = (size_t)mb << 20;
  }
  long long chunk = std::max<long long>(1, (long long)chunk_bytes / (long long)per_matrix);
  chunk = std::min(chunk, batch);
  if (layout == JPD_LAYOUT_SOA && chunk > 1) chunk &= ~1LL;  // keep planes 16-byte aligned

  for (int s = 0; s < 3; ++s) {
    Slot& sl = pipe.slot[s];
    if (!sl.stream) JPD_CK(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
    int rc;
    if ((rc = grow(&sl.in, &sl.in_b, sizeof(T) * nn * chunk))) return rc;
    if ((rc = grow(&sl.evals, &sl.evals_b, sizeof(T) * n * chunk))) return rc;
    if (calc && (rc = grow(&sl.evecs, &sl.evecs_b, sizeof(T) * nn * chunk))) return rc;
    if ((rc = grow((void**)&sl.iters, &sl.iters_b, sizeof(int) * chunk))) return rc;
  }

  int status = JPD_OK;
  long long done = 0;
  for (int i = 0; done < batch; ++i) {
    Slot& sl = pipe.slot[i % 3];
    const long long cnt = std::min(chunk, batch - done);
    cudaStream_t st = sl.stream;
    T* d_in = (T*)sl.in; T* d_ev = (T*)sl.evals; T* d_vec = calc ? (T*)sl.evecs : nullptr;
    if (layout == JPD_LAYOUT_AOS) {
      JPD_CK(cudaMemcpyAsync(d_in, h_mat + (size_t)done * nn, sizeof(T) * nn * cnt, cudaMemcpyHostToDevice, st));
    } else {
      JPD_CK(cudaMemcpy2DAsync(d_in, sizeof(T) * cnt, h_mat + done, sizeof(T) * host_pitch,
                               sizeof(T) * cnt, nn, cudaMemcpyHostToDevice, st));
    }
    status = diagonalize_device<T>(d_in, d_ev, d_vec, sl.iters, n, cnt, layout, sort, calc, sweeps, st);
    if (status != JPD_OK) break;
    if (layout == JPD_LAYOUT_AOS) {
      JPD_CK(cudaMemcpyAsync(h_evals + (size_t)done * n, d_ev, sizeof(T) * n * cnt, cudaMemcpyDeviceToHost, st));
      if (calc) JPD_CK(cudaMemcpyAsync(h_evecs + (size_t)done * nn, d_vec, sizeof(T) * nn * cnt, cudaMemcpyDeviceToHost, st));
    } else {
      JPD_CK(cudaMemcpy2DAsync(h_evals + done, sizeof(T) * host_pitch, d_ev, sizeof(T) * cnt,
                               sizeof(T) * cnt, n, cudaMemcpyDeviceToHost, st));
      if (calc) JPD_CK(cudaMemcpy2DAsync(h_evecs + done, sizeof(T) * host_pitch, d_vec, sizeof(T) * cnt,
                                         sizeof(T) * cnt, nn, cudaMemcpyDeviceToHost, st));
    }
    if (h_iters) JPD_CK(cudaMemcpyAsync(h_iters + done, sl.iters, sizeof(int) * cnt, cudaMemcpyDeviceToHost, st));
    done += cnt;
  }
  for (int s = 0; s < 3; ++s) {
    cudaError_t e = cudaStreamSynchronize(pipe.slot[s].stream);
    if (e != cudaSuccess && status == JPD_OK) {
      g_last_cuda_error = std::string("cudaStreamSynchronize: ") + cudaGetErrorString(e);
      status = JPD_ERR_CUDA;
    }
  }
  return status;
}

template <typename T>
int diagonalize_host_multi(const T* h_mat, T* h_evals, T* h_evecs, int* h_iters, int n,
                           long long batch, int layout, int sort, int calc, int sweeps,
                           int n_devices, const int* devices) {
  if (n_devices < 1 || n_devices > kMaxDevices) return JPD_ERR_INVALID_ARG;
  if (n < 1 || n > JPD_MAX_N || batch < 0) return JPD_ERR_INVALID_ARG;
  std::vector<int> rc(n_devices, JPD_OK);
  std::vector<std::string> msg(n_devices);
  std::vector<std::thread> workers;
  const size_t nn = (size_t)n * n;
  for (int g = 0; g < n_devices; ++g) {
    workers.emplace_back([&, g]() {
      long long first = 0, count = 0;
      jpd_slice(batch, n_devices, g, &first, &count);
      if (count == 0) return;
      const int dev = devices ? devices[g] : g;
      const bool soa = layout == JPD_LAYOUT_SOA;
      rc[g] = diagonalize_host<T>(h_mat + (soa ? (size_t)first : (size_t)first * nn),
                                  h_evals + (soa ? (size_t)first : (size_t)first * n),
                                  h_evecs ? h_evecs + (soa ? (size_t)first : (size_t)first * nn) : nullptr,
                                  h_iters ? h_iters + first : nullptr, n, count, batch, layout, sort,
                                  calc, sweeps, dev);
      if (rc[g] != JPD_OK) msg[g] = g_last_cuda_error;
    });
  }
  for (auto& w : workers) w.join();
  for (int g = 0; g < n_devices; ++g)
    if (rc[g] != JPD_OK) { g_last_cuda_error = msg[g]; return rc[g]; }
  return JPD_OK;
}

// ------------------------------------------------------------------ synthetic inputs
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  unsigned long long z = x + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ double u_pm1(unsigned long long seed, unsigned long long b, unsigned long long e) {
  const unsigned long long z = splitmix64(seed ^ (b * 65536ull + e));
  return (double)(z >> 11) * 0x1.0p-53 * 2.0 - 1.0;
}

template <typename T>
__global__ void jpd_synth_kernel(T* __restrict__ mat, int n, long long batch, int soa, int kind,
                                 unsigned long long seed, long long b0) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const unsigned long long gb = (unsigned long long)(b0 + b);
  auto put = [&](int i, int j, double v) {
    const size_t e = (size_t)i * n + j;
    mat[soa ? e * (size_t)batch + (size_t)b : (size_t)b * n * n + e] = (T)v;
  };
  if (kind == 1) {  // n == 4 checked by the host wrapper
    double R[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) R[e] = u_pm1(seed, gb, (unsigned long long)e);
    const double xx = R[0], xy = R[1], xz = R[2], yx = R[3], yy = R[4], yz = R[5], zx = R[6], zy = R[7], zz = R[8];
    const double s00 = __dadd_rn(__dadd_rn(xx, yy), zz), s01 = __dsub_rn(yz, zy), s02 = __dsub_rn(zx, xz), s03 = __dsub_rn(xy, yx);
    const double s11 = __dsub_rn(__dsub_rn(xx, yy), zz), s12 = __dadd_rn(xy, yx), s13 = __dadd_rn(zx, xz);
    const double s22 = __dsub_rn(__dsub_rn(yy, xx), zz), s23 = __dadd_rn(yz, zy);
    const double s33 = __dsub_rn(__dsub_rn(zz, xx), yy);
    put(0, 0, s00); put(0, 1, s01); put(0, 2, s02); put(0, 3, s03);
    put(1, 0, s01); put(1, 1, s11); put(1, 2, s12); put(1, 3, s13);
    put(2, 0, s02); put(2, 1, s12); put(2, 2, s22); put(2, 3, s23);
    put(3, 0, s03); put(3, 1, s13); put(3, 2, s23); put(3, 3, s33);
    return;
  }
  for (int i = 0; i < n; ++i)
    for (int j = i; j < n; ++j) {
      const double v = u_pm1(seed, gb, (unsigned long long)(i * n + j));
      put(i, j, v);
      if (j != i) put(j, i, v);
    }
}

template <typename T>
int fill_synthetic(T* d_mat, int n, long long batch, int layout, int kind, unsigned long long seed,
                   long long b0, cudaStream_t st) {
  if (n < 1 || n > 256 || batch < 0 || !d_mat) return JPD_ERR_INVALID_ARG;
  if (kind != 0 && kind != 1) return JPD_ERR_INVALID_ARG;
  if (kind == 1 && n != 4) return JPD_ERR_INVALID_ARG;
  if (layout != JPD_LAYOUT_AOS && layout != JPD_LAYOUT_SOA) return JPD_ERR_INVALID_ARG;
  if (batch == 0) return JPD_OK;
  const int threads = 128;
  const long long blocks = (batch + threads - 1) / threads;
  jpd_synth_kernel<T><<<dim3((unsigned)blocks), dim3(threads), 0, st>>>(d_mat, n, batch, layout == JPD_LAYOUT_SOA, kind, seed, b0);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  JPD_CK(cudaGetLastError());
  return JPD_OK;
}

// ------------------------------------------------------------------ FMA-pipe probe
// Dependent-chain-free FMA stream: 8 independent accumulators per thread, `iters` rounds.
// bench.py times it with CUDA events to get the measured FP64 / FP32 FMA peak that the
// FP-bound rooflines (n >= 16) are quoted against (MEASURED_PEAKS.json has no FP64 figure).
template <typename T>
__global__ void jpd_fma_probe_kernel(T* out, int iters, T a, T b) {
  T x0 = T(threadIdx.x), x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = jfma(x0, a, b); x1 = jfma(x1, a, b); x2 = jfma(x2, a, b); x3 = jfma(x3, a, b);
    x4 = jfma(x4, a, b); x5 = jfma(x5, a, b); x6 = jfma(x6, a, b); x7 = jfma(x7, a, b);
  }
  const T r = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (r == T(-12345.678)) out[0] = r;  // never true; keeps the chain alive
}

}  // namespace

// =================================================================== extern "C"
extern "C" {

int jpd_diagonalize_batched_f64(const double* d_mat, double* d_evals, double* d_evecs, int* d_n_iters,
                                int n, long long batch, int layout, int sort_criteria, int calc_evecs,
                                int max_num_sweeps, void* cuda_stream) {
  return diagonalize_device<double>(d_mat, d_evals, d_evecs, d_n_iters, n, batch, layout, sort_criteria,
                                    calc_evecs, max_num_sweeps, (cudaStream_t)cuda_stream);
}
int jpd_diagonalize_batched_f32(const float* d_mat, float* d_evals, float* d_evecs, int* d_n_iters,
                                int n, long long batch, int layout, int sort_criteria, int calc_evecs,
                                int max_num_sweeps, void* cuda_stream) {
  return diagonalize_device<float>(d_mat, d_evals, d_evecs, d_n_iters, n, batch, layout, sort_criteria,
                                   calc_evecs, max_num_sweeps, (cudaStream_t)cuda_stream);
}

int jpd_diagonalize_batched_host_f64(const double* h_mat, double* h_evals, double* h_evecs, int* h_n_iters,
                                     int n, long long batch, int layout, int sort_criteria, int calc_evecs,
                                     int max_num_sweeps, int device) {
  return diagonalize_host<double>(h_mat, h_evals, h_evecs, h_n_iters, n, batch, batch, layout,
                                  sort_criteria, calc_evecs ? 1 : 0, max_num_sweeps, device);
}
int jpd_diagonalize_batched_host_f32(const float* h_mat, float* h_evals, float* h_evecs, int* h_n_iters,
                                     int n, long long batch, int layout, int sort_criteria, int calc_evecs,
                                     int max_num_sweeps, int device) {
  return diagonalize_host<float>(h_mat, h_evals, h_evecs, h_n_iters, n, batch, batch, layout,
                                 sort_criteria, calc_evecs ? 1 : 0, max_num_sweeps, device);
}

int jpd_diagonalize_batched_host_multi_f64(const double* h_mat, double* h_evals, double* h_evecs,
                                           int* h_n_iters, int n, long long batch, int layout,
                                           int sort_criteria, int calc_evecs, int max_num_sweeps,
                                           int n_devices, const int* devices) {
  return diagonalize_host_multi<double>(h_mat, h_evals, h_evecs, h_n_iters, n, batch, layout,
                                        sort_criteria, calc_evecs ? 1 : 0, max_num_sweeps, n_devices, devices);
}
int jpd_diagonalize_batched_host_multi_f32(const float* h_mat, float* h_evals, float* h_evecs,
                                           int* h_n_iters, int n, long long batch, int layout,
                                           int sort_criteria, int calc_evecs, int max_num_sweeps,
                                           int n_devices, const int* devices) {
  return diagonalize_host_multi<float>(h_mat, h_evals, h_evecs, h_n_iters, n, batch, layout,
                                       sort_criteria, calc_evecs ? 1 : 0, max_num_sweeps, n_devices, devices);
}

int jpd_quaternion_from_correlation_f64(const double* d_corr, double* d_lambda, double* d_quat, int* d_n_iters,
                                        long long batch, int layout, int all_eigenpairs, int max_num_sweeps,
                                        void* cuda_stream) {
  return quaternion_device<double>(d_corr, d_lambda, d_quat, d_n_iters, batch, layout, all_eigenpairs,
                                   max_num_sweeps, (cudaStream_t)cuda_stream);
}
int jpd_quaternion_from_correlation_f32(const float* d_corr, float* d_lambda, float* d_quat, int* d_n_iters,
                                        long long batch, int layout, int all_eigenpairs, int max_num_sweeps,
                                        void* cuda_stream) {
  return quaternion_device<float>(d_corr, d_lambda, d_quat, d_n_iters, batch, layout, all_eigenpairs,
                                  max_num_sweeps, (cudaStream_t)cuda_stream);
}
int jpd_principal_axes_f64(const double* d_inertia6, double* d_moments, double* d_axes, int* d_n_iters,
                           long long batch, int layout, int max_num_sweeps, void* cuda_stream) {
  return principal_axes_device<double>(d_inertia6, d_moments, d_axes, d_n_iters, batch, layout,
                                       max_num_sweeps, (cudaStream_t)cuda_stream);
}
int jpd_principal_axes_f32(const float* d_inertia6, float* d_moments, float* d_axes, int* d_n_iters,
                           long long batch, int layout, int max_num_sweeps, void* cuda_stream) {
  return principal_axes_device<float>(d_inertia6, d_moments, d_axes, d_n_iters, batch, layout,
                                      max_num_sweeps, (cudaStream_t)cuda_stream);
}

int jpd_convert_layout_f64(const void* d_src, int src_layout, double* d_dst, int dst_layout, int n,
                           long long batch, void* cuda_stream) {
  return convert_layout<double>(d_src, src_layout, d_dst, dst_layout, n, batch, (cudaStream_t)cuda_stream);
}
int jpd_convert_layout_f32(const void* d_src, int src_layout, float* d_dst, int dst_layout, int n,
                           long long batch, void* cuda_stream) {
  return convert_layout<float>(d_src, src_layout, d_dst, dst_layout, n, batch, (cudaStream_t)cuda_stream);
}

void jpd_slice(long long batch, int n_parts, int part, long long* first, long long* count) {
  long long f = 0, c = 0;
  if (batch > 0 && n_parts > 0 && part >= 0 && part < n_parts) {
    const long long per = (batch + n_parts - 1) / n_parts;
    f = std::min(batch, (long long)part * per);
    c = std::min(per, batch - f);
  }
  if (first) *first = f;
  if (count) *count = c;
}

int jpd_fill_synthetic_f64(double* d_mat, int n, long long batch, int layout, int kind,
                           unsigned long long seed, long long b0, void* cuda_stream) {
  return fill_synthetic<double>(d_mat, n, batch, layout, kind, seed, b0, (cudaStream_t)cuda_stream);
}
int jpd_fill_synthetic_f32(float* d_mat, int n, long long batch, int layout, int kind,
                           unsigned long long seed, long long b0, void* cuda_stream) {
  return fill_synthetic<float>(d_mat, n, batch, layout, kind, seed, b0, (cudaStream_t)cuda_stream);
}

long long jpd_fma_probe(int elem_size, int blocks, int threads, int iters, void* d_scratch, void* cuda_stream) {
  if (blocks < 1 || threads < 32 || threads > 1024 || iters < 1 || !d_scratch) return JPD_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  if (elem_size == 8) jpd_fma_probe_kernel<double><<<blocks, threads, 0, st>>>((double*)d_scratch, iters, 0.999999, 1e-9);
  else if (elem_size == 4) jpd_fma_probe_kernel<float><<<blocks, threads, 0, st>>>((float*)d_scratch, iters, 0.999999f, 1e-9f);
  else return JPD_ERR_INVALID_ARG;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (cudaGetLastError() != cudaSuccess) return JPD_ERR_CUDA;
  return 8LL * iters * blocks * threads;  // FMAs issued
}

long long jpd_algorithmic_bytes(int n, int elem_size, int calc_evecs) {
  if (n < 1 || elem_size < 1) return -1;
  const long long nn = (long long)n * n;
  return (long long)elem_size * ((long long)n * (n + 1) / 2 + n + (calc_evecs ? nn : 0)) + 4;
}

int jpd_kernel_tier(int n, int elem_size, int calc_evecs) {
  if (n < 1 || n > JPD_MAX_N || (elem_size != 4 && elem_size != 8)) return JPD_ERR_INVALID_ARG;
  if (n <= kSmallMaxN) return 1;
  if (n <= kWarpMaxN) return 2;
  if (n <= kCtaMaxN) return 3;
  if (n <= kClusterMaxN) return 4;
  int limit = 232448, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) {
    DeviceInfo info;
    if (device_info(dev, &info) == JPD_OK) limit = info.smem_optin;
  } else {
    (void)cudaGetLastError();
  }
  const BlockPlan p = make_block_plan(n, elem_size, calc_evecs ? 1 : 0, limit);
  return 5;
}

long long jpd_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int jpd_release(void) {
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess) { (void)cudaGetLastError(); return JPD_OK; }
  int cur = 0;
  cudaGetDevice(&cur);
  for (int d = 0; d < n_dev && d < kMaxDevices; ++d) {
    Pipeline& pipe = g_pipe[d];
    std::lock_guard<std::mutex> lock(pipe.mutex);
    bool any = false;
    for (auto& sl : pipe.slot) any = any || sl.stream || sl.in || sl.evals || sl.evecs || sl.iters;
    any = any || pipe.small_pin || pipe.small_dev;
    if (!any) continue;
    cudaSetDevice(d);
    if (pipe.small_pin) { cudaFreeHost(pipe.small_pin); pipe.small_pin = nullptr; }
    if (pipe.small_dev) { cudaFree(pipe.small_dev); pipe.small_dev = nullptr; }
    for (auto& sl : pipe.slot) {
      if (sl.stream) { cudaStreamSynchronize(sl.stream); cudaStreamDestroy(sl.stream); }
      cudaFree(sl.in); cudaFree(sl.evals); cudaFree(sl.evecs); cudaFree(sl.iters);
      sl = Slot{};
    }
  }
  cudaSetDevice(cur);
  return JPD_OK;
}

const char* jpd_error_string(int code) {
  switch (code) {
    case JPD_OK: return "ok";
    case JPD_ERR_INVALID_ARG: return "invalid argument";
    case JPD_ERR_UNSUPPORTED: return "unsupported size or configuration";
    case JPD_ERR_CUDA: return "CUDA error (see jpd_last_cuda_error)";
    case JPD_ERR_ALLOC: return "allocation failed";
    default: return "unknown error";
  }
}
const char* jpd_last_cuda_error(void) { return g_last_cuda_error.c_str(); }
int jpd_version(void) { return 100; }

}  // extern "C"
