// Exercises the C++ front end (include/jacobi_pd_cuda.hpp) end to end on the GPU:
//  - jacobi_pd::Jacobi on the README example of the reference (README.md:58-99)
//  - jacobi_pd::BatchedJacobi<double>::DiagonalizeHost on a small AoS batch, checked by
//    residual / orthogonality / ordering;
//  - DiagonalizeHostPacked, QuaternionFromCorrelationHost, HostRegister against the plain entry.
// Exit code 0 = pass.  Run by tests/test_gpu_dropin.py.
#include <cmath>
#include <cstdio>
#include <vector>

#include "jacobi_pd.hpp"

static int fail(const char *what) {
  std::printf("FAIL: %s\n", what);
  return 1;
}

int main() {
  using namespace jacobi_pd;
  {  // README example: eigenvalues (3, 3, 0) in decreasing order
    int n = 3;
    double **M, **evecs;
    matrix_alloc_jpd::Alloc2D(n, n, &M);
    matrix_alloc_jpd::Alloc2D(n, n, &evecs);
    double evals[3];
    const double src[3][3] = {{2, 1, 1}, {1, 2, -1}, {1, -1, 2}};
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) M[i][j] = src[i][j];
    Jacobi<double, double *, double **, double const *const *> solver(n);
    int it = solver.Diagonalize(M, evals, evecs);
    if (it <= 0) return fail("README example did not converge");
    if (std::fabs(evals[0] - 3) > 1e-12 || std::fabs(evals[1] - 3) > 1e-12 || std::fabs(evals[2]) > 1e-12)
      return fail("README example eigenvalues");
    for (int k = 0; k < n; ++k)
      for (int a = 0; a < n; ++a) {
        double r = -evals[k] * evecs[k][a];
        for (int b = 0; b < n; ++b) r += src[a][b] * evecs[k][b];
        if (std::fabs(r) > 1e-12) return fail("README example residual");
      }
    matrix_alloc_jpd::Dealloc2D(&M);
    matrix_alloc_jpd::Dealloc2D(&evecs);
    if (M != nullptr || evecs != nullptr) return fail("Dealloc2D must null the pointer");
  }
  {  // batched host front end, n = 4
    const int n = 4;
    const long long B = 1000;
    std::vector<double> mat(B * n * n), evals(B * n), evecs(B * n * n);
    std::vector<int> iters(B);
    unsigned long long s = 88172645463325252ull;
    for (long long b = 0; b < B; ++b)
      for (int i = 0; i < n; ++i)
        for (int j = i; j < n; ++j) {
          s ^= s << 13; s ^= s >> 7; s ^= s << 17;
          double v = (double)(s >> 11) / 9007199254740992.0 * 2.0 - 1.0;
          mat[b * 16 + i * n + j] = mat[b * 16 + j * n + i] = v;
        }
    BatchedJacobi<double>::DiagonalizeHost(mat.data(), evals.data(), evecs.data(), iters.data(), n, B);
    for (long long b = 0; b < B; ++b) {
      if (iters[b] <= 0) return fail("batched: not converged");
      for (int k = 0; k + 1 < n; ++k)
        if (evals[b * n + k] < evals[b * n + k + 1]) return fail("batched: not sorted decreasing");
      for (int k = 0; k < n; ++k) {
        for (int a = 0; a < n; ++a) {
          double r = -evals[b * n + k] * evecs[b * 16 + k * n + a];
          for (int c = 0; c < n; ++c) r += mat[b * 16 + a * n + c] * evecs[b * 16 + k * n + c];
          if (std::fabs(r) > 4e-12 * 4) return fail("batched: residual");
        }
        for (int l = 0; l < n; ++l) {
          double d = (k == l) ? -1.0 : 0.0;
          for (int c = 0; c < n; ++c) d += evecs[b * 16 + k * n + c] * evecs[b * 16 + l * n + c];
          if (std::fabs(d) > 4e-12) return fail("batched: orthogonality");
        }
      }
    }
  }
  {  // packed upper-triangle host input and the fused colvars host entry: same results as the plain ones
    const int n = 4;
    const long long B = 500;
    std::vector<double> corr(B * 9), mat(B * 16), packed(B * 10), ev(B * 4), vec(B * 16), ev2(B * 4), vec2(B * 16);
    std::vector<double> lam(B), quat(B * 4);
    std::vector<int> iters(B);
    unsigned long long s = 1234567ull;
    for (long long b = 0; b < B; ++b) {
      double R[9];
      for (int e = 0; e < 9; ++e) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        R[e] = corr[b * 9 + e] = (double)(s >> 11) / 9007199254740992.0 * 2.0 - 1.0;
      }
      const double xx = R[0], xy = R[1], xz = R[2], yx = R[3], yy = R[4], yz = R[5], zx = R[6], zy = R[7], zz = R[8];
      const double S[4][4] = {{(xx + yy) + zz, yz - zy, zx - xz, xy - yx},
                              {0, (xx - yy) - zz, xy + yx, zx + xz},
                              {0, 0, (yy - xx) - zz, yz + zy},
                              {0, 0, 0, (zz - xx) - yy}};
      int t = 0;
      for (int i = 0; i < n; ++i)
        for (int j = i; j < n; ++j) { mat[b * 16 + i * n + j] = mat[b * 16 + j * n + i] = S[i][j]; packed[b * 10 + t++] = S[i][j]; }
    }
    BatchedJacobi<double>::HostRegister(mat.data(), mat.size() * sizeof(double));
    BatchedJacobi<double>::DiagonalizeHost(mat.data(), ev.data(), vec.data(), iters.data(), n, B);
    BatchedJacobi<double>::HostUnregister(mat.data());
    BatchedJacobi<double>::DiagonalizeHostPacked(packed.data(), ev2.data(), vec2.data(), iters.data(), n, B);
    if (ev != ev2 || vec != vec2) return fail("packed host input differs from the full input");
    BatchedJacobi<double>::QuaternionFromCorrelationHost(corr.data(), lam.data(), quat.data(), iters.data(), B);
    for (long long b = 0; b < B; ++b) {
      if (lam[b] != ev[b * 4]) return fail("fused host entry: leading eigenvalue differs");
      if (quat[b * 4] < 0) return fail("fused host entry: q0 < 0");
      for (int v = 0; v < 4; ++v)
        if (std::fabs(std::fabs(quat[b * 4 + v]) - std::fabs(vec[b * 16 + v])) != 0) return fail("fused host entry: quaternion differs");
    }
  }
  std::printf("test passed\n");
  return 0;
}
