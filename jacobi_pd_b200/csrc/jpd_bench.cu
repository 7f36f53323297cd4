// jpd_bench.cu -- bench / test scaffolding exported from libjpd.so but NOT part of the product
// ABI (declared in include/jpd_bench.h): the device twin of the oracle's synthetic input
// generator and the FMA-pipe probe that measures the FP64 / FP32 peak.
#include "jpd_internal.h"
#include "jpd_bench.h"
#include "jpd_common.cuh"

namespace {
using namespace jpd;
using namespace jpd_host;

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
  if (kind == 2) {  // n == 3 checked by the host wrapper: the correlation matrix itself (NOT symmetric)
#pragma unroll
    for (int e = 0; e < 9; ++e) mat[soa ? (size_t)e * (size_t)batch + (size_t)b : (size_t)b * 9 + e] = (T)u_pm1(seed, gb, (unsigned long long)e);
    return;
  }
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
  if (kind < 0 || kind > 2) return JPD_ERR_INVALID_ARG;
  if ((kind == 1 && n != 4) || (kind == 2 && n != 3)) return JPD_ERR_INVALID_ARG;
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

extern "C" {

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

}  // extern "C"
