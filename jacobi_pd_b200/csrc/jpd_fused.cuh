// jpd_fused.cuh -- the step BEFORE the path, fused into the tier-1 kernels for the two real
// callers of jacobi_pd (/root/reference/README.md:124-128 names them):
//
//  * colvars optimal rotation: the 4x4 quaternion-superposition overlap matrix S is built in
//    registers from the 3x3 correlation matrix C = sum x y^T (Horn 1987 / colvars
//    convention, the same formula as the synthetic generator kind 1), diagonalised
//    (Jacobi::Diagonalize, /root/reference/include/jacobi_pd.hpp:159-220, n = 4,
//    SORT_DECREASING_EVALS) and only the leading eigenpair -- the optimal-rotation
//    quaternion and its eigenvalue -- is written: 72 B in + 44 B out per matrix instead of
//    244 B.
//  * LAMMPS rigid bodies: the 3x3 inertia tensor arrives as its 6 unique components,
//    Diagonalize (n = 3, SORT_DECREASING_EVALS), then the post-step its callers apply:
//    eigenvectors as COLUMNS and a right-handed frame (ez flipped if (ex x ey).ez < 0).
//
// Both reuse small_solve() of jpd_small.cuh unchanged: same rotation, same skip test,
// same sweep logic, same return value (the producer is the `load` callback, so that the
// rare rescaling pass can rebuild the matrix from the inputs).
#pragma once
#include "jpd_small.cuh"

namespace jpd {

#if defined(__CUDACC__)

// S (upper triangle, tri order) from the row-major correlation matrix R[9]
template <typename T>
__device__ __forceinline__ void overlap_from_correlation(const T (&R)[9], T (&m)[10]) {
  const T xx = R[0], xy = R[1], xz = R[2], yx = R[3], yy = R[4], yz = R[5], zx = R[6], zy = R[7], zz = R[8];
  m[0] = (xx + yy) + zz;  m[1] = yz - zy;          m[2] = zx - xz;          m[3] = xy - yx;
  m[4] = (xx - yy) - zz;  m[5] = xy + yx;          m[6] = zx + xz;
  m[7] = (yy - xx) - zz;  m[8] = yz + zy;
  m[9] = (zz - xx) - yy;
}

// ALL = false: d_lambda[b], d_quat[4] (leading eigenpair, q0 >= 0)
// ALL = true : d_lambda = 4 eigenvalues (decreasing), d_quat = 4x4 eigenvector rows
template <typename T, bool SOA, bool ALL>
__global__ void __launch_bounds__(JPD_SMALL_THREADS, (SmallLaunch<4, T>::kMinBlocks))
jpd_quat_kernel(const T* __restrict__ corr, T* __restrict__ lambda, T* __restrict__ quat,
                int* __restrict__ n_iters, long long batch, int max_num_sweeps) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = gid < batch;
  const long long b = live ? gid : batch - 1;
  auto load = [&](T (&m)[10]) {
    T R[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = SOA ? __ldg(corr + (long long)k * batch + b) : __ldg(corr + b * 9 + k);
    overlap_from_correlation(R, m);
  };
  auto emit = [&](const T (&m)[10], const T (&e)[16], int iters) {
    T ev[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) ev[i] = m[SmallShape<4>::tri(i, i)];
    if (!live) return;
    if (n_iters) n_iters[b] = iters;
    if (ALL) {
      int pos[4];
      small_sort_positions<4, T>(ev, SORT_DECREASING_EVALS, pos);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (SOA) lambda[(long long)pos[r] * batch + b] = ev[r]; else lambda[b * 4 + pos[r]] = ev[r];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          if (SOA) quat[((long long)pos[r] * 4 + v) * batch + b] = e[r * 4 + v];
          else quat[b * 16 + pos[r] * 4 + v] = e[r * 4 + v];
        }
      }
    } else {
      // first maximum, as position 0 of the reference's decreasing selection sort (:483-487)
      T best = ev[0];
      int imax = 0;
#pragma unroll
      for (int i = 1; i < 4; ++i) { const bool gt = ev[i] > best; best = gt ? ev[i] : best; imax = gt ? i : imax; }
      T q[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        q[v] = e[v];
#pragma unroll
        for (int r = 1; r < 4; ++r) q[v] = (imax == r) ? e[r * 4 + v] : q[v];
      }
      const bool flip = q[0] < T(0);
#pragma unroll
      for (int v = 0; v < 4; ++v) q[v] = flip ? -q[v] : q[v];
      lambda[b] = best;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        if (SOA) quat[(long long)v * batch + b] = q[v]; else quat[b * 4 + v] = q[v];
      }
    }
  };
  small_solve<4, T, true>(max_num_sweeps, load, emit);
}

// inertia6 = (Ixx, Iyy, Izz, Iyz, Ixz, Ixy); moments decreasing; axes[i*3 + k] = component i
// of principal axis k (eigenvectors as columns), right-handed.
template <typename T, bool SOA>
__global__ void __launch_bounds__(JPD_SMALL_THREADS, (SmallLaunch<3, T>::kMinBlocks))
jpd_inertia_kernel(const T* __restrict__ inertia6, T* __restrict__ moments, T* __restrict__ axes,
                   int* __restrict__ n_iters, long long batch, int max_num_sweeps) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = gid < batch;
  const long long b = live ? gid : batch - 1;
  auto load = [&](T (&m)[6]) {
    T I[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) I[k] = SOA ? __ldg(inertia6 + (long long)k * batch + b) : __ldg(inertia6 + b * 6 + k);
    m[0] = I[0]; m[1] = I[5]; m[2] = I[4];   // (0,0) (0,1) (0,2)
    m[3] = I[1]; m[4] = I[3];                // (1,1) (1,2)
    m[5] = I[2];                             // (2,2)
  };
  auto emit = [&](const T (&m)[6], const T (&e)[9], int iters) {
    T ev[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) ev[i] = m[SmallShape<3>::tri(i, i)];
    // Destinations are permuted, not data (no gather through selects or local memory):
    // source row r is principal axis pos[r].  Right-handed frame: ez <- -ez if
    // (ex x ey) . ez < 0, i.e. if det(sorted rows) = det(e) * sign(permutation) < 0.
    int pos[3];
    small_sort_positions<3, T>(ev, SORT_DECREASING_EVALS, pos);
    const T cx = e[1] * e[5] - e[2] * e[4];
    const T cy = e[2] * e[3] - e[0] * e[5];
    const T cz = e[0] * e[4] - e[1] * e[3];
    const bool det_neg = (cx * e[6] + cy * e[7] + cz * e[8]) < T(0);
    const bool odd = ((pos[0] > pos[1]) != (pos[0] > pos[2])) != (pos[1] > pos[2]);
    const bool flip = det_neg != odd;
    if (!live) return;
    if (n_iters) n_iters[b] = iters;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      if (SOA) moments[(long long)pos[r] * batch + b] = ev[r]; else moments[b * 3 + pos[r]] = ev[r];
      const bool neg = flip && pos[r] == 2;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const T val = neg ? -e[r * 3 + i] : e[r * 3 + i];
        if (SOA) axes[((long long)(i * 3) + pos[r]) * batch + b] = val;
        else axes[b * 9 + i * 3 + pos[r]] = val;
      }
    }
  };
  small_solve<3, T, true>(max_num_sweeps, load, emit);
}

#endif  // __CUDACC__

}  // namespace jpd
