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

#include "jpd_internal.h"

namespace jpd_host {

namespace {
thread_local std::string g_last_cuda_error;
DeviceInfo g_dev[kMaxDevices];
std::mutex g_dev_mutex;
}  // namespace

std::atomic<long long> g_launches{0};

void set_last_cuda_error(const char* expr, cudaError_t e) {
  g_last_cuda_error = std::string(expr) + ": " + cudaGetErrorString(e);
}

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

void keep_freed_workspace_cached(int dev) {
  if (dev < 0 || dev >= kMaxDevices) return;
  std::lock_guard<std::mutex> lock(g_dev_mutex);
  if (g_dev[dev].pool_tuned) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  g_dev[dev].pool_tuned = true;
}

const char* last_cuda_error_cstr() { return g_last_cuda_error.c_str(); }

}  // namespace jpd_host

namespace {

using namespace jpd_host;

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
  if (n <= kSmallMaxN) return launch_tier1<T>(mat, evals, evecs, iters, n, batch, layout, sort, calc, sweeps, st);
  if (n <= kWarpMaxN) return launch_tier2<T>(mat, evals, evecs, iters, n, batch, layout, sort, calc, sweeps, st);
  if (n <= kCtaMaxN) return launch_tier3<T>(mat, evals, evecs, iters, n, batch, layout, sort, calc, sweeps, st);
  if (n <= kClusterMaxN) return launch_tier4<T>(mat, evals, evecs, iters, n, batch, layout, sort, calc, sweeps, st);
  return launch_tier5<T>(mat, evals, evecs, iters, n, batch, layout, sort, calc, sweeps, st);
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
  (void)calc_evecs;
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
const char* jpd_last_cuda_error(void) { return jpd_host::last_cuda_error_cstr(); }
int jpd_version(void) { return 100; }

}  // extern "C"
