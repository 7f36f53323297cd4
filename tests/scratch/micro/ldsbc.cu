// microbenchmark (run under ncu): shared-memory wavefronts of broadcast loads by width / alignment
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double* out, int iters, int off) {
  __shared__ __align__(32) double sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
  __syncthreads();
  double acc = 0;
  const int lane = threadIdx.x & 31;
  for (int it = 0; it < iters; ++it) {
    const int base = ((it * 10) & 1023) + off;     // warp-uniform
    if (MODE == 0) { double2 v = *reinterpret_cast<double2*>(&sm[2 * base]); acc += v.x + v.y; }                 // LDS.128 broadcast
    if (MODE == 1) { acc += sm[2 * base] + sm[2 * base + 1]; }                                                   // 2 x LDS.64 broadcast
    if (MODE == 2) { double2 v = *reinterpret_cast<double2*>(&sm[2 * (base + (lane >> 4))]); acc += v.x + v.y; } // 2 addresses (half warps)
    if (MODE == 3) { double2 v = *reinterpret_cast<double2*>(&sm[2 * (base + (lane >> 3))]); acc += v.x + v.y; } // 4 addresses (quarter warps)
    if (MODE == 4) { double2 v = *reinterpret_cast<double2*>(&sm[2 * (base + lane)]); acc += v.x + v.y; }        // 32 addresses, contiguous
    if (MODE == 5) { double2 v = *reinterpret_cast<double2*>(&sm[2 * (base + (lane & 7))]); acc += v.x + v.y; }  // 8 addresses, each by 4 lanes (one per quarter)
    if (MODE == 6) { acc += sm[2 * base]; }                                                                      // LDS.64 broadcast
    if (MODE == 7) { acc += sm[2 * (base + (lane & 15))]; }                                                      // LDS.64, 16 addresses x2
  }
  if (acc == 1.2345) out[threadIdx.x] = acc;
}
int main() {
  double* d; cudaMalloc(&d, 8192);
  k<0><<<1, 32>>>(d, 1000, 0); k<1><<<1, 32>>>(d, 1000, 0); k<2><<<1, 32>>>(d, 1000, 0); k<3><<<1, 32>>>(d, 1000, 0);
  k<4><<<1, 32>>>(d, 1000, 0); k<5><<<1, 32>>>(d, 1000, 0); k<6><<<1, 32>>>(d, 1000, 0); k<7><<<1, 32>>>(d, 1000, 0);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
