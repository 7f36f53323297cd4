// microbenchmark: throughput of the plane-rotation update (x,y) -> (c x - s y, s x + c y) on W
// register-resident pairs, by instruction form/order.  6 warps per scheduler, sm_100a.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int W = 8;

// FORM 0: per pair  m1 = s*y; x' = fma(c,x,-m1); m2 = c*y; y' = fma(s,x,m2)      (source order of the kernel)
// FORM 1: grouped   all m1 = s*y | all x' = fma(c,x,-m1) | all m2 = c*y | all y' = fma(s,x,m2)
// FORM 2: t-form    u = fma(-t,y,x); v = fma(t,x,y); x' = c*u; y' = c*v   (t = s/c)
// FORM 3: t-form grouped
// FORM 4: unnormalised (fast rotation): x' = fma(-t,y,x); y' = fma(t,x,y)   (2 instructions per pair)
template <int FORM, int FILL>
__global__ void __launch_bounds__(128, 6) k(double* out, int iters, double c0, double s0) {
  double x[W], y[W];
#pragma unroll
  for (int i = 0; i < W; ++i) { x[i] = 1.0 + threadIdx.x * 1e-3 + i; y[i] = 2.0 - threadIdx.x * 1e-3 + i; }
  double c = c0, s = s0 + threadIdx.x * 1e-9, t = s0 * 1.0001;
  unsigned q = threadIdx.x;
  for (int it = 0; it < iters; ++it) {
    // keep (c,s,t) loop-carried so nothing folds (3 extra FP64 instructions per W pairs)
    c = c * 0.9999999; s = s * 1.0000001; t = t * 1.0000001;
    if (FILL) {   // integer-pipe filler between updates, as the real sweep has
#pragma unroll
      for (int f = 0; f < FILL; ++f) q = (q ^ (q >> 3)) + 0x9e3779b9u;
    }
    if (FORM == 0) {
#pragma unroll
      for (int i = 0; i < W; ++i) {
        const double m1 = s * y[i]; const double xn = __fma_rn(c, x[i], -m1);
        const double m2 = c * y[i]; y[i] = __fma_rn(s, x[i], m2); x[i] = xn;
      }
    } else if (FORM == 1) {
      double m1[W], m2[W], xn[W];
#pragma unroll
      for (int i = 0; i < W; ++i) m1[i] = s * y[i];
#pragma unroll
      for (int i = 0; i < W; ++i) xn[i] = __fma_rn(c, x[i], -m1[i]);
#pragma unroll
      for (int i = 0; i < W; ++i) m2[i] = c * y[i];
#pragma unroll
      for (int i = 0; i < W; ++i) { y[i] = __fma_rn(s, x[i], m2[i]); x[i] = xn[i]; }
    } else if (FORM == 2) {
#pragma unroll
      for (int i = 0; i < W; ++i) {
        const double u = __fma_rn(-t, y[i], x[i]); const double v = __fma_rn(t, x[i], y[i]);
        x[i] = c * u; y[i] = c * v;
      }
    } else if (FORM == 3) {
      double u[W], v[W];
#pragma unroll
      for (int i = 0; i < W; ++i) u[i] = __fma_rn(-t, y[i], x[i]);
#pragma unroll
      for (int i = 0; i < W; ++i) v[i] = __fma_rn(t, x[i], y[i]);
#pragma unroll
      for (int i = 0; i < W; ++i) x[i] = c * u[i];
#pragma unroll
      for (int i = 0; i < W; ++i) y[i] = c * v[i];
    } else {
#pragma unroll
      for (int i = 0; i < W; ++i) {
        const double u = __fma_rn(-t, y[i], x[i]); y[i] = __fma_rn(t, x[i], y[i]); x[i] = u;
      }
    }
  }
  double r = c + s + t + q;
#pragma unroll
  for (int i = 0; i < W; ++i) r += x[i] + y[i];
  if (r == 12345.678) out[threadIdx.x] = r;
}

template <int FORM, int FILL> void run(double* d, const char* name) {
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<FORM, FILL><<<148 * 6, 128>>>(d, 10, 0.8, 0.6);
  cudaEventRecord(e0);
  k<FORM, FILL><<<148 * 6, 128>>>(d, iters, 0.8, 0.6);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double cyc = ms * 1e-3 * 1.965e9 / iters / 6.0;
  const int nfp = (FORM == 4 ? 2 : 4) * W + 3;
  printf("%-34s fill %2d: %7.1f cycles per %d pairs per warp = %5.2f per FP64 instruction (%d)\n", name, FILL, cyc, W, cyc / nfp, nfp);
}

int main() {
  double* d; cudaMalloc(&d, 1024 * 8);
  run<0, 0>(d, "form 0 mul,fma,mul,fma per pair");
  run<1, 0>(d, "form 1 grouped");
  run<2, 0>(d, "form 2 t-form per pair");
  run<3, 0>(d, "form 3 t-form grouped");
  run<4, 0>(d, "form 4 unnormalised 2 fma");
  run<0, 8>(d, "form 0 mul,fma,mul,fma per pair");
  run<1, 8>(d, "form 1 grouped");
  run<0, 16>(d, "form 0 mul,fma,mul,fma per pair");
  run<1, 16>(d, "form 1 grouped");
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
