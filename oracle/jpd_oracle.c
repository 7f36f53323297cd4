/* jpd_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library (oracle/libjpd_oracle.so).  The product
 * path (jacobi_pd_b200/libjpd.so, include/jacobi_pd_cuda.hpp) never calls it and
 * has no CPU fallback.
 *
 * What it is: a plain-C restatement of the serial classical-Jacobi algorithm of
 *   /root/reference/include/jacobi_pd.hpp   (Jacobi<...>::Diagonalize, :159-220)
 * for Scalar = double and Scalar = float (body in jpd_oracle_body.inc, each
 * function citing the reference lines it follows), plus the host twin of the
 * counter-based synthetic-input generator that the CUDA library implements on
 * the device (jpdo_fill_synthetic_*).
 *
 * Pinning: PINNED.  tests/test_oracle_pinned.py checks this restatement
 *   (1) against the reference's own known answers (README.md:67-69 example,
 *       SURVEY 8(c) behaviours) stored in tests/golden/known_answers.json,
 *   (2) bit-for-bit against fixtures produced by the UNMODIFIED reference header
 *       (tests/golden/ref_fixtures_*.npz, made by tests/golden/make_fixtures.py
 *       through oracle/_ref/libjpd_ref.so), and
 *   (3) bit-for-bit against oracle/_ref/libjpd_ref.so itself when it is present.
 *
 * Build: oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).  -ffp-contract=off
 * matters: the reference's x86-64 -O2 build has no FMA, and bit parity needs the
 * same roundings.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define S double
#define SFX f64
#define FABS fabs
#define SQRT sqrt
#include "jpd_oracle_body.inc"
#undef S
#undef SFX
#undef FABS
#undef SQRT

#define S float
#define SFX f32
#define FABS fabsf
#define SQRT sqrtf
#include "jpd_oracle_body.inc"
#undef S
#undef SFX
#undef FABS
#undef SQRT

int jpdo_host_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ---- synthetic inputs: host twin of jpd_fill_synthetic_* (csrc/jpd_synth.cu) ----
 * u(seed,b,e) = splitmix64(seed ^ (b*65536 + e)) -> top 53 bits -> U(-1,1).
 * kind 0: upper-triangle entries e=i*n+j (i<=j) iid U(-1,1), mirrored   
 * kind 1: n==4 only, 3x3 correlation matrix R (e=0..8) iid U(-1,1) folded into the
 *         traceless quaternion-superposition overlap matrix (Horn 1987)  
 * layout 0 (AoS): m[b*n*n + i*n + j]; layout 1 (SoA-interleaved): m[(i*n+j)*batch + b].
 * Matrices b0 .. b0+batch-1 of the global stream are written at local index 0..batch-1.
 */
static inline uint64_t splitmix64(uint64_t x) {
  uint64_t z = x + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static inline double u_pm1(uint64_t seed, uint64_t b, uint64_t e) {
  uint64_t z = splitmix64(seed ^ (b * 65536ull + e));
  return (double)(z >> 11) * 0x1.0p-53 * 2.0 - 1.0;
}

static void synth_matrix(double *m /* n*n row-major */, int n, int kind, uint64_t seed, uint64_t b) {
  if (kind == 1 && n == 4) {
    double R[9];
    for (int e = 0; e < 9; ++e) R[e] = u_pm1(seed, b, (uint64_t)e);
    const double xx = R[0], xy = R[1], xz = R[2], yx = R[3], yy = R[4], yz = R[5],
                 zx = R[6], zy = R[7], zz = R[8];
    m[0] = (xx + yy) + zz; m[1] = yz - zy;        m[2] = zx - xz;          m[3] = xy - yx;
    m[5] = (xx - yy) - zz; m[6] = xy + yx;        m[7] = zx + xz;
    m[10] = (yy - xx) - zz; m[11] = yz + zy;
    m[15] = (zz - xx) - yy;
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < i; ++j) m[i * 4 + j] = m[j * 4 + i];
    return;
  }
  for (int i = 0; i < n; ++i)
    for (int j = i; j < n; ++j) {
      double v = u_pm1(seed, b, (uint64_t)(i * n + j));
      m[i * n + j] = v;
      m[j * n + i] = v;
    }
}

#define DEFINE_FILL(SFX, S)                                                                  \
  int jpdo_fill_synthetic_##SFX(S *mat, int n, long long batch, int layout, int kind,        \
                                unsigned long long seed, long long b0) {                     \
    if (n <= 0 || n > 256 || batch < 0 || (kind == 1 && n != 4)) return -1;                   \
    _Pragma("omp parallel")                                                                  \
    {                                                                                        \
      double *tmp = (double *)malloc(sizeof(double) * (size_t)n * n);                        \
      _Pragma("omp for schedule(static)")                                                    \
      for (long long b = 0; b < batch; ++b) {                                                \
        synth_matrix(tmp, n, kind, (uint64_t)seed, (uint64_t)(b0 + b));                      \
        for (int e = 0; e < n * n; ++e) {                                                    \
          size_t at = layout ? (size_t)e * (size_t)batch + (size_t)b                         \
                             : (size_t)b * n * n + (size_t)e;                                \
          mat[at] = (S)tmp[e];                                                               \
        }                                                                                    \
      }                                                                                      \
      free(tmp);                                                                             \
    }                                                                                        \
    return 0;                                                                                \
  }
DEFINE_FILL(f64, double)
DEFINE_FILL(f32, float)
