// microbenchmark: FP64 pipe throughput by operand pattern (register-file read bandwidth?) and
// dependent-issue latencies, sm_100a.  6 warps per scheduler unless stated.
#include <cstdio>
#include <cuda_runtime.h>

// MODE 0: d[i] = fma(a[i%8], b[(i+3)%8], d[i])      three distinct changing register operands
// MODE 1: d[i] = a[i%8] * b[(i+3)%8]                two distinct operands (DMUL), d overwritten (then folded)
// MODE 2: d[i] = fma(d[i], a0, b0)                  two operands constant (reuse cache)
// MODE 3: d[i] = fma(a[i%8], b0, d[i])              one constant
// MODE 4: d[i] = a[i%8] + d[i]                      DADD
// MODE 5: d[i] = d[i] * a[i%8]                      DMUL accumulating
template <int MODE>
__global__ void __launch_bounds__(128, 6) thr(double* out, int iters, double s0) {
  double a[8], b[8], d[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = 1.0 + (threadIdx.x + i) * 1e-9 * s0; b[i] = 1.0 - (threadIdx.x + 2 * i) * 1e-9 * s0; }
#pragma unroll
  for (int i = 0; i < 16; ++i) d[i] = i * s0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int k = i % 16;
      if (MODE == 0) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[k]) : "d"(a[i % 8]), "d"(b[(i + 3) % 8]));
      if (MODE == 1) asm volatile("mul.rn.f64 %0, %1, %2;" : "=d"(d[k]) : "d"(a[i % 8]), "d"(b[(i + 3) % 8]));
      if (MODE == 2) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[k]) : "d"(a[0]), "d"(b[0]));
      if (MODE == 3) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[k]) : "d"(a[i % 8]), "d"(b[0]));
      if (MODE == 4) asm volatile("add.rn.f64 %0, %1, %0;" : "+d"(d[k]) : "d"(a[i % 8]));
      if (MODE == 5) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(d[k]) : "d"(a[i % 8]));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += d[i];
  if (s == 12345.678) out[threadIdx.x] = s;
}

// dependent chain, one warp per scheduler: latency per instruction
template <int MODE>
__global__ void __launch_bounds__(128, 1) lat(double* out, int iters, double s0) {
  double x = 1.0 + threadIdx.x * 1e-12 * s0, a = 1.0 + 1e-9 * s0, b = 1e-12 * s0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if (MODE == 0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x) : "d"(a), "d"(b));
      if (MODE == 1) asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(x));
      if (MODE == 2) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(x) : "d"(a));
      if (MODE == 3) { asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(x)); asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x) : "d"(a), "d"(b)); }
    }
  }
  if (x == 12345.678) out[threadIdx.x] = x;
}

template <int MODE> void run_thr(double* d, const char* name) {
  const int iters = 4000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  thr<MODE><<<148 * 6, 128>>>(d, 10, 1.0);
  cudaEventRecord(e0);
  thr<MODE><<<148 * 6, 128>>>(d, iters, 1.0);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("throughput %-44s %6.2f cycles per warp instruction per scheduler\n", name, ms * 1e-3 * 1.965e9 / iters / 6.0 / 32.0);
}
template <int MODE> void run_lat(double* d, const char* name, int per) {
  const int iters = 4000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  lat<MODE><<<148, 128>>>(d, 10, 1.0);
  cudaEventRecord(e0);
  lat<MODE><<<148, 128>>>(d, iters, 1.0);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("latency    %-44s %6.2f cycles per dependent step\n", name, ms * 1e-3 * 1.965e9 / iters / 32.0);
}

int main() {
  double* d; cudaMalloc(&d, 1024 * 8);
  run_thr<0>(d, "DFMA d=a_i*b_j+d (3 distinct regs)");
  run_thr<1>(d, "DMUL d=a_i*b_j (2 distinct regs)");
  run_thr<2>(d, "DFMA d=d*a0+b0 (2 constant regs)");
  run_thr<3>(d, "DFMA d=a_i*b0+d (1 constant reg)");
  run_thr<4>(d, "DADD d=a_i+d");
  run_thr<5>(d, "DMUL d=d*a_i");
  run_lat<0>(d, "DFMA chain", 1);
  run_lat<2>(d, "DMUL chain", 1);
  run_lat<1>(d, "MUFU.RSQ64H chain", 1);
  run_lat<3>(d, "MUFU.RSQ64H + DFMA chain (per pair)", 2);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
