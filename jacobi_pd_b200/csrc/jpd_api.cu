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

// smallest n that goes to the block-Jacobi tier (default: everything beyond the cluster tier);
// JPD_BLOCK_MIN_N (> 128) moves the boundary down -- tuning aid for the tier-4 / tier-5 crossover
int block_tier_min_n() {
  static int v = -1;
  if (v < 0) {
    v = kClusterMaxN + 1;
    if (const char* env = std::getenv("JPD_BLOCK_MIN_N")) {
      const long x = std::strtol(env, nullptr, 10);
      if (x > kCtaMaxN && x <= kClusterMaxN + 1) v = (int)x;
    }
  }
  return v;
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
  if (n <= kSmallMaxN) return launch_tier1<T>(mat, evals, evecs, iters, n, batch, layout, sort, calc, sweeps, st);
  if (n <= kWarpMaxN) return launch_tier2<T>(mat, evals, evecs, iters, n, batch, layout, sort, calc, sweeps, st);
  if (n <= kCtaMaxN) return launch_tier3<T>(mat, evals, evecs, iters, n, batch, layout, sort, calc, sweeps, st);
  if (n <= kClusterMaxN && n < block_tier_min_n())
    return launch_tier4<T>(mat, evals, evecs, iters, n, batch, layout, sort, calc, sweeps, st);
  return launch_tier5<T>(mat, evals, evecs, iters, n, batch, layout, sort, calc, sweeps, st);
}

// ------------------------------------------------------------------ host pipeline
// Generic for every host entry: a call is a set of host arrays ("spans": inputs, outputs, and
// device-only scratch), each with a fixed number of elements per matrix, laid out AoS
// (matrix-major) or SoA (element-major planes of `host_pitch` matrices).  The batch is cut
// into chunks of ~64 MB that go through three slots (stream + device buffers): H2D of chunk
// k+1, the kernels of chunk k and D2H of chunk k-1 overlap.  Per-device staging is grown on
// demand and kept until jpd_release().  One host call at a time per device (mutex).
constexpr int kMaxSpans = 5;
struct Span {
  const void* h_in = nullptr;   // host source (uploaded)          -- at most one of the two;
  void* h_out = nullptr;        // host destination (downloaded)   -- neither: device scratch
  size_t per = 0;               // elements per matrix
  size_t elem = 0;              // bytes per element
  bool soa = false;             // element-major planes (pitch = host_pitch matrices)
};
struct Slot {
  cudaStream_t stream = nullptr;
  void* buf[kMaxSpans] = {};
  size_t cap[kMaxSpans] = {};
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

// Makes `device` current for the duration of a host entry and restores the caller's device on
// every return path (a host application may have another device current).
struct DeviceGuard {
  int prev = -1;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) { (void)cudaGetLastError(); prev = -1; }
    if (prev != device) err = cudaSetDevice(device);
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int grow(void** p, size_t* have, size_t want) {
  if (*have >= want) return JPD_OK;
  if (*p) JPD_CK(cudaFree(*p));
  *p = nullptr; *have = 0;
  JPD_CK(cudaMalloc(p, want));
  *have = want;
  return JPD_OK;
}

inline size_t up16(size_t x) { return (x + 15) / 16 * 16; }

// launch(d_bufs, count, stream): enqueue the kernels of one chunk; d_bufs[k] is the device
// image of span k for `count` matrices (AoS: contiguous; SoA: planes of `count`).
template <typename Launch>
int host_pipeline(int device, long long batch, long long host_pitch, const Span* spans, int n_spans,
                  Launch launch) {
  if (device < 0 || device >= kMaxDevices || n_spans < 1 || n_spans > kMaxSpans) return JPD_ERR_INVALID_ARG;
  if (batch == 0) return JPD_OK;
  DeviceGuard guard(device);
  JPD_CK(guard.err);
  Pipeline& pipe = g_pipe[device];
  std::lock_guard<std::mutex> lock(pipe.mutex);

  size_t per_matrix = 0;
  for (int k = 0; k < n_spans; ++k) per_matrix += spans[k].per * spans[k].elem;

  // ---- small calls (the drop-in scalar class: batch = 1): everything through ONE pinned
  // staging buffer -- one H2D copy of the inputs, the kernels, one D2H copy of the outputs,
  // one synchronisation -- instead of the chunked three-stream pipeline
  if (host_pitch == batch) {
    size_t off[kMaxSpans + 1];
    size_t in_end = 0, out_begin = 0, pos = 0;
    bool ordered = true;   // inputs first, then outputs/scratch
    for (int pass = 0; pass < 2; ++pass)
      for (int k = 0; k < n_spans; ++k) {
        const bool is_in = spans[k].h_in != nullptr;
        if ((pass == 0) != is_in) continue;
        off[k] = pos;
        pos = up16(pos + spans[k].per * spans[k].elem * (size_t)batch);
        if (pass == 0) in_end = pos;
      }
    out_begin = in_end;
    (void)ordered;
    if (pos <= kSmallCallBytes) {
      if (!pipe.small_pin) {
        JPD_CK(cudaHostAlloc(&pipe.small_pin, kSmallCallBytes, cudaHostAllocDefault));
        JPD_CK(cudaMalloc(&pipe.small_dev, kSmallCallBytes));
      }
      if (!pipe.slot[0].stream) JPD_CK(cudaStreamCreateWithFlags(&pipe.slot[0].stream, cudaStreamNonBlocking));
      cudaStream_t st = pipe.slot[0].stream;
      char* hp = static_cast<char*>(pipe.small_pin);
      char* dp = static_cast<char*>(pipe.small_dev);
      void* d_bufs[kMaxSpans];
      for (int k = 0; k < n_spans; ++k) {
        d_bufs[k] = dp + off[k];
        if (spans[k].h_in) std::memcpy(hp + off[k], spans[k].h_in, spans[k].per * spans[k].elem * (size_t)batch);
      }
      int status = JPD_OK;
      cudaError_t e = in_end ? cudaMemcpyAsync(dp, hp, in_end, cudaMemcpyHostToDevice, st) : cudaSuccess;
      if (e == cudaSuccess) {
        status = launch(d_bufs, batch, st);
        if (status == JPD_OK && pos > out_begin)
          e = cudaMemcpyAsync(hp + out_begin, dp + out_begin, pos - out_begin, cudaMemcpyDeviceToHost, st);
      }
      const cudaError_t es = cudaStreamSynchronize(st);   // also on errors: nothing may stay in flight
      if (e == cudaSuccess) e = es;
      if (e != cudaSuccess && status == JPD_OK) { set_last_cuda_error("small host call", e); status = JPD_ERR_CUDA; }
      if (status != JPD_OK) return status;
      for (int k = 0; k < n_spans; ++k)
        if (spans[k].h_out) std::memcpy(spans[k].h_out, hp + off[k], spans[k].per * spans[k].elem * (size_t)batch);
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
  if (chunk > 1) chunk &= ~1LL;  // keep SoA planes 16-byte aligned

  for (int s = 0; s < 3; ++s) {
    Slot& sl = pipe.slot[s];
    if (!sl.stream) JPD_CK(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
    for (int k = 0; k < n_spans; ++k) {
      const int rc = grow(&sl.buf[k], &sl.cap[k], spans[k].per * spans[k].elem * (size_t)chunk);
      if (rc != JPD_OK) return rc;
    }
  }

  // From here on an error must not return before all three streams are idle: copies into the
  // caller's buffers may still be in flight, and the slots are reused by the next call.
  int status = JPD_OK;
  cudaError_t err = cudaSuccess;
  const char* what = "";
  auto copy = [&](const Span& sp, void* dev, long long done, long long cnt, bool up, cudaStream_t st) {
    if (err != cudaSuccess) return;
    const cudaMemcpyKind kind = up ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    if (!sp.soa) {
      char* h = (char*)(up ? const_cast<void*>(sp.h_in) : sp.h_out) + (size_t)done * sp.per * sp.elem;
      err = up ? cudaMemcpyAsync(dev, h, sp.per * sp.elem * (size_t)cnt, kind, st)
               : cudaMemcpyAsync(h, dev, sp.per * sp.elem * (size_t)cnt, kind, st);
    } else {
      char* h = (char*)(up ? const_cast<void*>(sp.h_in) : sp.h_out) + (size_t)done * sp.elem;
      const size_t hp = sp.elem * (size_t)host_pitch, dpitch = sp.elem * (size_t)cnt;
      err = up ? cudaMemcpy2DAsync(dev, dpitch, h, hp, dpitch, sp.per, kind, st)
               : cudaMemcpy2DAsync(h, hp, dev, dpitch, dpitch, sp.per, kind, st);
    }
    if (err != cudaSuccess) what = up ? "host-to-device copy" : "device-to-host copy";
  };
  long long done = 0;
  for (int i = 0; done < batch && status == JPD_OK && err == cudaSuccess; ++i) {
    Slot& sl = pipe.slot[i % 3];
    const long long cnt = std::min(chunk, batch - done);
    for (int k = 0; k < n_spans; ++k)
      if (spans[k].h_in) copy(spans[k], sl.buf[k], done, cnt, true, sl.stream);
    if (err != cudaSuccess) break;
    status = launch(sl.buf, cnt, sl.stream);
    if (status != JPD_OK) break;
    for (int k = 0; k < n_spans; ++k)
      if (spans[k].h_out) copy(spans[k], sl.buf[k], done, cnt, false, sl.stream);
    done += cnt;
  }
  for (int s = 0; s < 3; ++s) {
    const cudaError_t e = cudaStreamSynchronize(pipe.slot[s].stream);
    if (e != cudaSuccess && err == cudaSuccess) { err = e; what = "cudaStreamSynchronize"; }
  }
  if (err != cudaSuccess && status == JPD_OK) { set_last_cuda_error(what, err); status = JPD_ERR_CUDA; }
  return status;
}

// host_pitch: matrices between consecutive SoA planes in the HOST arrays (== the caller's
// full batch, also when only a slice [first, first+batch) is processed).
template <typename T>
int diagonalize_host(const T* h_mat, T* h_evals, T* h_evecs, int* h_iters, int n, long long batch,
                     long long host_pitch, int layout, int sort, int calc, int sweeps, int device) {
  if (n < 1 || n > JPD_MAX_N || batch < 0) return JPD_ERR_INVALID_ARG;
  if (layout != JPD_LAYOUT_AOS && layout != JPD_LAYOUT_SOA) return JPD_ERR_INVALID_ARG;
  if (batch == 0) return JPD_OK;
  if (!h_mat || !h_evals || (calc && !h_evecs)) return JPD_ERR_INVALID_ARG;
  const size_t nn = (size_t)n * n;
  const bool soa = layout == JPD_LAYOUT_SOA;
  Span sp[4];
  sp[0].h_in = h_mat;    sp[0].per = nn;             sp[0].elem = sizeof(T);   sp[0].soa = soa;
  sp[1].h_out = h_evals; sp[1].per = (size_t)n;      sp[1].elem = sizeof(T);   sp[1].soa = soa;
  sp[2].h_out = calc ? h_evecs : nullptr; sp[2].per = calc ? nn : 0; sp[2].elem = sizeof(T); sp[2].soa = soa;
  sp[3].h_out = h_iters; sp[3].per = 1;              sp[3].elem = sizeof(int); sp[3].soa = false;  // scratch if NULL
  return host_pipeline(device, batch, host_pitch, sp, 4, [&](void* const* d, long long cnt, cudaStream_t st) {
    return diagonalize_device<T>((const T*)d[0], (T*)d[1], calc ? (T*)d[2] : nullptr, (int*)d[3], n, cnt, layout,
                                 sort, calc, sweeps, st);
  });
}

// Packed upper triangle in (n(n+1)/2 scalars per matrix instead of n*n: what the reference
// actually reads, jacobi_pd.hpp:167-171); unpacked on the device in front of the solver.
template <typename T>
int diagonalize_host_packed(const T* h_upper, T* h_evals, T* h_evecs, int* h_iters, int n, long long batch,
                            int layout, int sort, int calc, int sweeps, int device) {
  if (n < 1 || n > JPD_MAX_N || batch < 0) return JPD_ERR_INVALID_ARG;
  if (layout != JPD_LAYOUT_AOS && layout != JPD_LAYOUT_SOA) return JPD_ERR_INVALID_ARG;
  if (batch == 0) return JPD_OK;
  if (!h_upper || !h_evals || (calc && !h_evecs)) return JPD_ERR_INVALID_ARG;
  const size_t nn = (size_t)n * n, tri = (size_t)n * (n + 1) / 2;
  const bool soa = layout == JPD_LAYOUT_SOA;
  Span sp[5];
  sp[0].h_in = h_upper;  sp[0].per = tri;        sp[0].elem = sizeof(T);   sp[0].soa = soa;
  sp[1].h_out = h_evals; sp[1].per = (size_t)n;  sp[1].elem = sizeof(T);   sp[1].soa = soa;
  sp[2].h_out = calc ? h_evecs : nullptr; sp[2].per = calc ? nn : 0; sp[2].elem = sizeof(T); sp[2].soa = soa;
  sp[3].h_out = h_iters; sp[3].per = 1;          sp[3].elem = sizeof(int); sp[3].soa = false;
  sp[4].per = nn;        sp[4].elem = sizeof(T);                                              // full matrices (scratch)
  return host_pipeline(device, batch, batch, sp, 5, [&](void* const* d, long long cnt, cudaStream_t st) {
    int rc = convert_layout<T>(d[0], soa ? 3 : 2, (T*)d[4], soa ? 1 : 0, n, cnt, st);
    if (rc != JPD_OK) return rc;
    return diagonalize_device<T>((const T*)d[4], (T*)d[1], calc ? (T*)d[2] : nullptr, (int*)d[3], n, cnt, layout,
                                 sort, calc, sweeps, st);
  });
}

// Fused producers with host buffers: 72 B up + 44 B down per matrix (fp64, leading eigenpair)
// for colvars instead of 128 + 168 B through jpd_diagonalize_batched_host_*.
template <typename T>
int quaternion_host(const T* h_corr, T* h_lambda, T* h_quat, int* h_iters, long long batch, int layout, int all,
                    int sweeps, int device) {
  if (batch < 0 || (layout != JPD_LAYOUT_AOS && layout != JPD_LAYOUT_SOA)) return JPD_ERR_INVALID_ARG;
  if (batch == 0) return JPD_OK;
  if (!h_corr || !h_lambda || !h_quat) return JPD_ERR_INVALID_ARG;
  const bool soa = layout == JPD_LAYOUT_SOA;
  Span sp[4];
  sp[0].h_in = h_corr;    sp[0].per = 9;             sp[0].elem = sizeof(T);   sp[0].soa = soa;
  sp[1].h_out = h_lambda; sp[1].per = all ? 4 : 1;   sp[1].elem = sizeof(T);   sp[1].soa = soa;
  sp[2].h_out = h_quat;   sp[2].per = all ? 16 : 4;  sp[2].elem = sizeof(T);   sp[2].soa = soa;
  sp[3].h_out = h_iters;  sp[3].per = 1;             sp[3].elem = sizeof(int); sp[3].soa = false;
  return host_pipeline(device, batch, batch, sp, 4, [&](void* const* d, long long cnt, cudaStream_t st) {
    return quaternion_device<T>((const T*)d[0], (T*)d[1], (T*)d[2], (int*)d[3], cnt, layout, all, sweeps, st);
  });
}

template <typename T>
int principal_axes_host(const T* h_inertia6, T* h_moments, T* h_axes, int* h_iters, long long batch, int layout,
                        int sweeps, int device) {
  if (batch < 0 || (layout != JPD_LAYOUT_AOS && layout != JPD_LAYOUT_SOA)) return JPD_ERR_INVALID_ARG;
  if (batch == 0) return JPD_OK;
  if (!h_inertia6 || !h_moments || !h_axes) return JPD_ERR_INVALID_ARG;
  const bool soa = layout == JPD_LAYOUT_SOA;
  Span sp[4];
  sp[0].h_in = h_inertia6; sp[0].per = 6; sp[0].elem = sizeof(T);   sp[0].soa = soa;
  sp[1].h_out = h_moments; sp[1].per = 3; sp[1].elem = sizeof(T);   sp[1].soa = soa;
  sp[2].h_out = h_axes;    sp[2].per = 9; sp[2].elem = sizeof(T);   sp[2].soa = soa;
  sp[3].h_out = h_iters;   sp[3].per = 1; sp[3].elem = sizeof(int); sp[3].soa = false;
  return host_pipeline(device, batch, batch, sp, 4, [&](void* const* d, long long cnt, cudaStream_t st) {
    return principal_axes_device<T>((const T*)d[0], (T*)d[1], (T*)d[2], (int*)d[3], cnt, layout, sweeps, st);
  });
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

int jpd_diagonalize_batched_host_packed_f64(const double* h_upper, double* h_evals, double* h_evecs, int* h_n_iters,
                                            int n, long long batch, int layout, int sort_criteria, int calc_evecs,
                                            int max_num_sweeps, int device) {
  return diagonalize_host_packed<double>(h_upper, h_evals, h_evecs, h_n_iters, n, batch, layout, sort_criteria,
                                         calc_evecs ? 1 : 0, max_num_sweeps < 0 ? 0 : max_num_sweeps, device);
}
int jpd_diagonalize_batched_host_packed_f32(const float* h_upper, float* h_evals, float* h_evecs, int* h_n_iters,
                                            int n, long long batch, int layout, int sort_criteria, int calc_evecs,
                                            int max_num_sweeps, int device) {
  return diagonalize_host_packed<float>(h_upper, h_evals, h_evecs, h_n_iters, n, batch, layout, sort_criteria,
                                        calc_evecs ? 1 : 0, max_num_sweeps < 0 ? 0 : max_num_sweeps, device);
}
int jpd_quaternion_from_correlation_host_f64(const double* h_corr, double* h_lambda, double* h_quat, int* h_n_iters,
                                             long long batch, int layout, int all_eigenpairs, int max_num_sweeps,
                                             int device) {
  return quaternion_host<double>(h_corr, h_lambda, h_quat, h_n_iters, batch, layout, all_eigenpairs ? 1 : 0,
                                 max_num_sweeps, device);
}
int jpd_quaternion_from_correlation_host_f32(const float* h_corr, float* h_lambda, float* h_quat, int* h_n_iters,
                                             long long batch, int layout, int all_eigenpairs, int max_num_sweeps,
                                             int device) {
  return quaternion_host<float>(h_corr, h_lambda, h_quat, h_n_iters, batch, layout, all_eigenpairs ? 1 : 0,
                                max_num_sweeps, device);
}
int jpd_principal_axes_host_f64(const double* h_inertia6, double* h_moments, double* h_axes, int* h_n_iters,
                                long long batch, int layout, int max_num_sweeps, int device) {
  return principal_axes_host<double>(h_inertia6, h_moments, h_axes, h_n_iters, batch, layout, max_num_sweeps, device);
}
int jpd_principal_axes_host_f32(const float* h_inertia6, float* h_moments, float* h_axes, int* h_n_iters,
                                long long batch, int layout, int max_num_sweeps, int device) {
  return principal_axes_host<float>(h_inertia6, h_moments, h_axes, h_n_iters, batch, layout, max_num_sweeps, device);
}

int jpd_host_register(void* h_ptr, unsigned long long bytes) {
  if (!h_ptr || bytes == 0) return JPD_ERR_INVALID_ARG;
  JPD_CK(cudaHostRegister(h_ptr, (size_t)bytes, cudaHostRegisterPortable));
  return JPD_OK;
}
int jpd_host_unregister(void* h_ptr) {
  if (!h_ptr) return JPD_ERR_INVALID_ARG;
  JPD_CK(cudaHostUnregister(h_ptr));
  return JPD_OK;
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
    for (auto& sl : pipe.slot) {
      any = any || sl.stream;
      for (void* p : sl.buf) any = any || p;
    }
    any = any || pipe.small_pin || pipe.small_dev;
    if (!any) continue;
    cudaSetDevice(d);
    if (pipe.small_pin) { cudaFreeHost(pipe.small_pin); pipe.small_pin = nullptr; }
    if (pipe.small_dev) { cudaFree(pipe.small_dev); pipe.small_dev = nullptr; }
    for (auto& sl : pipe.slot) {
      if (sl.stream) { cudaStreamSynchronize(sl.stream); cudaStreamDestroy(sl.stream); }
      for (void* p : sl.buf) cudaFree(p);
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
