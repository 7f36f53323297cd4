// microbenchmark: cost of cluster barriers (4 CTAs x 512 threads), cycles per barrier
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __cluster_dims__(4,1,1) __launch_bounds__(512,1) k(long long* out, int iters) {
  extern __shared__ double sm[];
  sm[threadIdx.x] = threadIdx.x;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
    if (MODE == 1) { asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory"); }
    if (MODE == 2) { __syncthreads(); }
    if (MODE == 3) { __syncthreads(); asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory"); }
    if (MODE == 4) { if (threadIdx.x < 128) asm volatile("fence.acq_rel.cluster;" ::: "memory"); __syncthreads(); asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory"); }
    if (MODE == 5) { // remote store by 128 threads + fence + light barrier
      if (threadIdx.x < 128) { unsigned a = (unsigned)__cvta_generic_to_shared(&sm[threadIdx.x]); unsigned r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(threadIdx.x & 3));
        asm volatile("st.shared::cluster.f64 [%0], %1;" :: "r"(r), "d"((double)i) : "memory"); asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
      __syncthreads(); asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory"); }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}
int main() {
  long long* d; cudaMalloc(&d, 8);
  const int iters = 20000;
  const char* names[] = {"release/acquire", "relaxed", "__syncthreads", "syncthreads+relaxed", "fence(128thr)+sync+relaxed", "remote st+fence+sync+relaxed"};
  #define RUN(M) { cudaFuncSetAttribute(k<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, 180000); k<M><<<4*33, 512, 180000>>>(d, iters); cudaError_t e = cudaDeviceSynchronize(); long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); printf("%-32s %8.1f cycles/barrier (%s)\n", names[M], (double)h / iters, cudaGetErrorString(e)); }
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5)
  return 0;
}
