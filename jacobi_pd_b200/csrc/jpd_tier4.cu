// jpd_tier4.cu -- see jpd_internal.h; part of libjpd.so (no CPU path: every entry ends in a kernel launch).
#include <algorithm>
#include <mutex>

#include "jpd_internal.h"
#include "jpd_cluster.cuh"
#include "jpd_cluster4.cuh"
#include <cstdlib>

namespace jpd_host {
using namespace jpd;

namespace {
// ------------------------------------------------------------------ tier 4 launch (4-CTA cluster per matrix)
static_assert(kClusterMaxN == ClusterShape::kN, "tier-4 bound");

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

// block-ordering kernel on the cluster (jpd_cluster4.cuh)
template <typename T>
int launch_cluster4(const T* mat, T* evals, T* evecs, int* iters, int n, long long batch, int layout,
                    int sort, int calc, int sweeps, cudaStream_t st) {
  using Sh = Cl4Shape;
  const size_t smem = Sh::block_bytes(sizeof(T));
  auto kernel = calc ? jpd_cluster4_kernel<T, true> : jpd_cluster4_kernel<T, false>;
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
// development switches: JPD_TIER4_BLOCK=1 forces the block-ordering kernel at every size, =0 the first-generation one
int tier4_switch() {
  static const int v = [] { const char* e = std::getenv("JPD_TIER4_BLOCK"); return e ? (e[0] == '1' ? 1 : (e[0] == '0' ? 0 : -1)) : -1; }();
  return v;
}

}  // namespace

template <typename T>
int launch_tier4(JPD_TIER_ARGS(T)) {
  // Measured on B200 (tests/scratch/t4sweep.py; DESIGN.md "tier 4"): since its pivots travel as bulk copies
  // (jpd_cluster4.cuh) the block-ordering kernel is ahead at every size and in both precisions -- 264 matrices,
  // fp64: n = 129 14.3 ms against 17.2 ms, n = 256 36.3 against 50.3; fp32: 9.1 against 14.5, 23.8 against 39.3.
  // The first-generation kernel stays for JPD_TIER4_BLOCK=0 (tests run both).
  const int sw = tier4_switch();
  const bool block_order = sw != 0;
  if (block_order) return launch_cluster4<T>(mat, evals, evecs, iters, n, batch, layout, sort, calc, sweeps, st);
  return launch_cluster<T>(mat, evals, evecs, iters, n, batch, layout, sort, calc, sweeps, st);
}

template int launch_tier4<double>(JPD_TIER_ARGS(double));
template int launch_tier4<float>(JPD_TIER_ARGS(float));

}  // namespace jpd_host
