// Exercises the DEVICE-pointer entry points of the C++ front end (include/jacobi_pd_cuda.hpp) from a
// CUDA C++ caller: BatchedJacobi<double>::Diagonalize (SoA-interleaved, the native layout),
// ::QuaternionFromCorrelation and ::PrincipalAxes (the fused producers for colvars / LAMMPS).
// Checks: residual, orthogonality, ordering, agreement of the fused quaternion path with the explicit
// 4x4 path.  Exit code 0 = pass.  Built by tests/dropin/Makefile with nvcc, run by tests/test_gpu_dropin.py.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <vector>

#include "jacobi_pd_cuda.hpp"

static int fail(const char *what) {
  std::printf("FAIL: %s\n", what);
  return 1;
}
#define CK(x) do { if ((x) != cudaSuccess) return fail(#x); } while (0)

int main() {
  using jacobi_pd::BatchedJacobi;
  using jacobi_pd::Layout;
  const long long B = 4096;
  unsigned long long s = 88172645463325252ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)(s >> 11) / 9007199254740992.0 * 2.0 - 1.0; };

  // ---- colvars: correlation matrices -> overlap matrices on the host (reference formula) ----
  std::vector<double> corr(B * 9), S(16 * B);   // S in SoA-interleaved layout: S[(i*4+j)*B + b]
  for (long long b = 0; b < B; ++b) {
    double R[9];
    for (int k = 0; k < 9; ++k) R[k] = corr[b * 9 + k] = rnd();
    const double xx = R[0], xy = R[1], xz = R[2], yx = R[3], yy = R[4], yz = R[5], zx = R[6], zy = R[7], zz = R[8];
    const double M[4][4] = {{(xx + yy) + zz, yz - zy, zx - xz, xy - yx},
                            {yz - zy, (xx - yy) - zz, xy + yx, zx + xz},
                            {zx - xz, xy + yx, (yy - xx) - zz, yz + zy},
                            {xy - yx, zx + xz, yz + zy, (zz - xx) - yy}};
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) S[(size_t)(i * 4 + j) * B + b] = M[i][j];
  }
  double *d_corr, *d_S, *d_ev, *d_vec, *d_lam, *d_q;
  int *d_it;
  CK(cudaMalloc(&d_corr, sizeof(double) * 9 * B));
  CK(cudaMalloc(&d_S, sizeof(double) * 16 * B));
  CK(cudaMalloc(&d_ev, sizeof(double) * 4 * B));
  CK(cudaMalloc(&d_vec, sizeof(double) * 16 * B));
  CK(cudaMalloc(&d_lam, sizeof(double) * B));
  CK(cudaMalloc(&d_q, sizeof(double) * 4 * B));
  CK(cudaMalloc(&d_it, sizeof(int) * B));
  CK(cudaMemcpy(d_corr, corr.data(), sizeof(double) * 9 * B, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_S, S.data(), sizeof(double) * 16 * B, cudaMemcpyHostToDevice));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  try {
    BatchedJacobi<double>::Diagonalize(d_S, d_ev, d_vec, d_it, 4, B, Layout::SoAInterleaved, JPD_SORT_DECREASING_EVALS, true, 50, st);
    BatchedJacobi<double>::QuaternionFromCorrelation(d_corr, d_lam, d_q, nullptr, B, Layout::AoS, false, 50, st);
  } catch (const std::exception &e) {
    return fail(e.what());
  }
  CK(cudaStreamSynchronize(st));
  std::vector<double> ev(4 * B), vec(16 * B), lam(B), q(4 * B);
  std::vector<int> it(B);
  CK(cudaMemcpy(ev.data(), d_ev, sizeof(double) * 4 * B, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(vec.data(), d_vec, sizeof(double) * 16 * B, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(lam.data(), d_lam, sizeof(double) * B, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(q.data(), d_q, sizeof(double) * 4 * B, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(it.data(), d_it, sizeof(int) * B, cudaMemcpyDeviceToHost));
  for (long long b = 0; b < B; ++b) {
    if (it[b] <= 0) return fail("device batch: not converged");
    for (int k = 0; k + 1 < 4; ++k)
      if (ev[(size_t)k * B + b] < ev[(size_t)(k + 1) * B + b]) return fail("device batch: not sorted decreasing");
    for (int k = 0; k < 4; ++k)
      for (int a = 0; a < 4; ++a) {
        double r = -ev[(size_t)k * B + b] * vec[(size_t)(k * 4 + a) * B + b];
        for (int c = 0; c < 4; ++c) r += S[(size_t)(a * 4 + c) * B + b] * vec[(size_t)(k * 4 + c) * B + b];
        if (std::fabs(r) > 4e-12 * 4) return fail("device batch: residual");
      }
    // fused path = leading eigenpair of the explicit path, sign fixed by q0 >= 0
    if (lam[b] != ev[b]) return fail("fused quaternion: eigenvalue differs from the explicit path");
    const double sgn = vec[(size_t)0 * B + b] < 0 ? -1.0 : 1.0;
    for (int v = 0; v < 4; ++v)
      if (q[b * 4 + v] != sgn * vec[(size_t)v * B + b]) return fail("fused quaternion: eigenvector differs from the explicit path");
    if (q[b * 4] < 0) return fail("fused quaternion: q0 < 0");
  }

  // ---- LAMMPS: inertia tensors -> principal axes ----
  std::vector<double> I6(6 * B), mom(3 * B), axes(9 * B);
  for (long long b = 0; b < B; ++b) {
    double I[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int k = 0; k < 8; ++k) {
      const double x = rnd(), y = 2 * rnd(), z = 3 * rnd(), r2 = x * x + y * y + z * z, p[3] = {x, y, z};
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) I[i][j] += (i == j ? r2 : 0.0) - p[i] * p[j];
    }
    const double v6[6] = {I[0][0], I[1][1], I[2][2], I[1][2], I[0][2], I[0][1]};
    for (int k = 0; k < 6; ++k) I6[b * 6 + k] = v6[k];
  }
  double *d_I6, *d_mom, *d_axes;
  CK(cudaMalloc(&d_I6, sizeof(double) * 6 * B));
  CK(cudaMalloc(&d_mom, sizeof(double) * 3 * B));
  CK(cudaMalloc(&d_axes, sizeof(double) * 9 * B));
  CK(cudaMemcpy(d_I6, I6.data(), sizeof(double) * 6 * B, cudaMemcpyHostToDevice));
  try {
    BatchedJacobi<double>::PrincipalAxes(d_I6, d_mom, d_axes, d_it, B, Layout::AoS, 50, st);
  } catch (const std::exception &e) {
    return fail(e.what());
  }
  CK(cudaStreamSynchronize(st));
  CK(cudaMemcpy(mom.data(), d_mom, sizeof(double) * 3 * B, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(axes.data(), d_axes, sizeof(double) * 9 * B, cudaMemcpyDeviceToHost));
  for (long long b = 0; b < B; ++b) {
    const double *v6 = &I6[b * 6], *A = &axes[b * 9];
    const double I[3][3] = {{v6[0], v6[5], v6[4]}, {v6[5], v6[1], v6[3]}, {v6[4], v6[3], v6[2]}};
    if (mom[b * 3] < mom[b * 3 + 1] || mom[b * 3 + 1] < mom[b * 3 + 2]) return fail("principal axes: moments not decreasing");
    for (int k = 0; k < 3; ++k)
      for (int a = 0; a < 3; ++a) {
        double r = -mom[b * 3 + k] * A[a * 3 + k];          // column k is axis k
        for (int c = 0; c < 3; ++c) r += I[a][c] * A[c * 3 + k];
        if (std::fabs(r) > 1e-11 * mom[b * 3]) return fail("principal axes: residual");
      }
    const double det = A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) + A[2] * (A[3] * A[7] - A[4] * A[6]);
    if (std::fabs(det - 1.0) > 1e-11) return fail("principal axes: frame is not right-handed");
  }
  jpd_release();
  std::printf("test passed\n");
  return 0;
}
