// microbenchmark: does an FP64 instruction cost one or two issue slots on sm_100a?
// F DFMAs (8 independent chains) interleaved with O other instructions per loop body, at the
// tier-1 kernel's occupancy (6 warps per scheduler).  Prints cycles per body per scheduler-warp.
#include <cstdio>
#include <cuda_runtime.h>

template <int F, int O, int KIND>
__global__ void __launch_bounds__(128, 6) k(double* out, int iters, double a, double b, unsigned q) {
  double ch[8];
  unsigned iv[8];
  float fv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { ch[i] = threadIdx.x * 1e-3 + i; iv[i] = threadIdx.x * 7u + i; fv[i] = threadIdx.x + i; }
  const float fa = (float)a, fb = (float)b;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < F; ++j) {
      asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(ch[j % 8]) : "d"(a), "d"(b));
      const int n_other = ((j + 1) * O) / F - (j * O) / F;
#pragma unroll
      for (int o = 0; o < n_other; ++o) {
        const int slot = (j * O / F + o) % 8;
        if (KIND == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(iv[slot]) : "r"(q), "r"(iv[(slot + 1) % 8]));
        if (KIND == 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(fv[slot]) : "f"(fa), "f"(fb));
        if (KIND == 2) { if ((j + o) & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(iv[slot]) : "r"(q), "r"(iv[(slot + 1) % 8]));
                         else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(fv[slot]) : "f"(fa), "f"(fb)); }
        if (KIND == 3) asm volatile("{.reg .pred p; setp.ne.u32 p, %2, 0; selp.f64 %0, %0, %1, p;}" : "+d"(ch[(slot + 4) % 8]) : "d"(a), "r"(q));
      }
    }
  }
  double s = 0; unsigned t = 0; float f = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { s += ch[i]; t ^= iv[i]; f += fv[i]; }
  if (s == 12345.678 || t == 77u || f == 3.25f) out[threadIdx.x] = s + t + f;
}

template <int F, int O, int KIND>
void run(double* d, const char* kind) {
  const int iters = 2000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<F, O, KIND><<<148 * 6, 128>>>(d, 10, 1.0000001, 1e-9, 3u);
  cudaEventRecord(e0);
  k<F, O, KIND><<<148 * 6, 128>>>(d, iters, 1.0000001, 1e-9, 3u);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  // 6 warps per scheduler; cycles per body per warp at 1965 MHz
  const double cyc = ms * 1e-3 * 1.965e9 / iters / 6.0;
  printf("%-8s F=%3d O=%3d  cycles/body/warp %7.1f   2F=%d  2F+O=%d  max(2F,F+O)=%d\n", kind, F, O, cyc, 2 * F, 2 * F + O,
         (2 * F > F + O ? 2 * F : F + O));
}

int main() {
  double* d; cudaMalloc(&d, 1024 * 8);
#define ROW(K, NAME) run<32, 0, K>(d, NAME); run<32, 8, K>(d, NAME); run<32, 16, K>(d, NAME); run<32, 24, K>(d, NAME); run<32, 32, K>(d, NAME); run<32, 48, K>(d, NAME); run<32, 64, K>(d, NAME);
  ROW(0, "lop3") ROW(1, "ffma") ROW(2, "mixed")
  run<32, 16, 3>(d, "selp64"); run<32, 32, 3>(d, "selp64");
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
