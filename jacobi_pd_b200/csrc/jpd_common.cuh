// jpd_common.cuh -- scalar traits and the Jacobi rotation used by every kernel tier.
//
// Everything here is __host__ __device__ so that tests/host_sim can run the exact
// per-matrix logic on the CPU (logic check only -- the product has no CPU path).
//
// Reference semantics being reproduced (algebra, not code):
//   CalcRot            /root/reference/include/jacobi_pd.hpp:233-256
//   convergence test   /root/reference/include/jacobi_pd.hpp:191
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>

#if defined(__CUDACC__)
#define JPD_HD __host__ __device__ __forceinline__
#define JPD_D __device__ __forceinline__
#define JPD_CX __host__ __device__
#else
#define JPD_HD inline
#define JPD_D inline
#define JPD_CX
#endif

namespace jpd {

enum SortCriteria : int {  // values as jacobi_pd.hpp:56-62
  DO_NOT_SORT = 0,
  SORT_DECREASING_EVALS = 1,
  SORT_INCREASING_EVALS = 2,
  SORT_DECREASING_ABS_EVALS = 3,
  SORT_INCREASING_ABS_EVALS = 4
};

template <typename T> struct Fp;
template <> struct Fp<double> {
  static constexpr int kMantBits = 52, kBias = 1023, kExpMask = 0x7ff, kSafe = 480;
};
template <> struct Fp<float> {
  static constexpr int kMantBits = 23, kBias = 127, kExpMask = 0xff, kSafe = 50;
};

// biased exponent field of x (0 = zero/denormal, kExpMask = inf/nan); integer pipe.
JPD_HD int expo_field(double x) {
#if defined(__CUDA_ARCH__)
  return (__double2hiint(x) >> 20) & 0x7ff;
#else
  std::uint64_t u; std::memcpy(&u, &x, 8);
  return (int)((u >> 52) & 0x7ff);
#endif
}
JPD_HD int expo_field(float x) {
#if defined(__CUDA_ARCH__)
  return (__float_as_int(x) >> 23) & 0xff;
#else
  std::uint32_t u; std::memcpy(&u, &x, 4);
  return (int)((u >> 23) & 0xff);
#endif
}
JPD_HD bool sign_bit(double x) {
#if defined(__CUDA_ARCH__)
  return __double2hiint(x) < 0;
#else
  return std::signbit(x);
#endif
}
JPD_HD bool sign_bit(float x) {
#if defined(__CUDA_ARCH__)
  return __float_as_int(x) < 0;
#else
  return std::signbit(x);
#endif
}
// x == +-0, decided on the integer pipe (a DSETP would cost an FP64 issue slot)
JPD_HD bool is_zero(double x) {
#if defined(__CUDA_ARCH__)
  return (((unsigned)__double2hiint(x) << 1) | (unsigned)__double2loint(x)) == 0u;
#else
  return x == 0.0;
#endif
}
JPD_HD bool is_zero(float x) {
#if defined(__CUDA_ARCH__)
  return ((unsigned)__float_as_int(x) << 1) == 0u;
#else
  return x == 0.0f;
#endif
}
// v with its sign flipped iff the sign bit of `src` is set (integer pipe: one LOP3 on the
// high word instead of an FP64 negate plus two selects)
JPD_HD double xor_sign(double v, double src) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(__double2hiint(v) ^ (__double2hiint(src) & (int)0x80000000), __double2loint(v));
#else
  return std::signbit(src) ? -v : v;
#endif
}
JPD_HD float xor_sign(float v, float src) {
#if defined(__CUDA_ARCH__)
  return __int_as_float(__float_as_int(v) ^ (__float_as_int(src) & (int)0x80000000));
#else
  return std::signbit(src) ? -v : v;
#endif
}
// 2^(field - bias) built on the integer pipe; field must be in [1, kExpMask-1].
JPD_HD double pow2_from_field(int field, double) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(field << 20, 0);
#else
  std::uint64_t u = (std::uint64_t)field << 52; double r; std::memcpy(&r, &u, 8); return r;
#endif
}
JPD_HD float pow2_from_field(int field, float) {
#if defined(__CUDA_ARCH__)
  return __int_as_float(field << 23);
#else
  std::uint32_t u = (std::uint32_t)field << 23; float r; std::memcpy(&r, &u, 4); return r;
#endif
}

JPD_HD double jfma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return std::fma(a, b, c);
#endif
}
JPD_HD float jfma(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
  return __fmaf_rn(a, b, c);
#else
  return std::fma(a, b, c);
#endif
}
JPD_HD double jabs(double a) { return fabs(a); }
JPD_HD float jabs(float a) { return fabsf(a); }

// 1/sqrt(x) to ~1 ulp for normal x: one SFU seed (MUFU.RSQ64H / MUFU.RSQ) plus one
// correction on the FMA pipe.  fp64: cubic step, 5 DFMA-class ops, no division, no
// IEEE sqrt expansion (the reference's 3 div + 2 sqrt expand to ~100 FP64-pipe
// instructions on sm_100; see DESIGN.md "rotation").
// (split in seed + correction so that callers can put independent work between the two)
JPD_HD double rsqrt_seed(double x) {
#if defined(__CUDA_ARCH__)
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
  return y0;
#else
  return 1.0 / std::sqrt(x);
#endif
}
#if defined(__CUDACC__)
// 3/8 as a constant-bank operand: an FP64 instruction takes one immediate, and a literal
// here makes ptxas re-materialise the second constant with two MOVs per use (r1 SASS: ~7
// of the 76 instructions of a rotation); c[bank][offset] operands cost nothing.
static __constant__ double jpd_c_three_eighths = 0.375;
static __constant__ double jpd_c_half = 0.5;
#endif
JPD_HD double rsqrt_correct(double x, double y0) {
#if defined(__CUDA_ARCH__)
  double xy = x * y0;
  double e = __fma_rn(-xy, y0, 1.0);       // 1 - x*y0^2, |e| <~ 2^-21
#if defined(JPD_NO_CONST_BANK)
  double r = __fma_rn(0.375, e, 0.5);      // 1/2 + 3/8 e
#else
  double r = __fma_rn(jpd_c_three_eighths, e, 0.5);
#endif
  double ey = e * y0;
  return __fma_rn(r, ey, y0);              // y0 (1 + e/2 + 3/8 e^2)
#else
  (void)x;
  return y0;
#endif
}
JPD_HD float rsqrt_seed(float x) {
#if defined(__CUDA_ARCH__)
  float y0;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(x));
  return y0;
#else
  return (float)(1.0 / std::sqrt((double)x));
#endif
}
JPD_HD float rsqrt_correct(float x, float y0) {
#if defined(__CUDA_ARCH__)
  float xy = x * y0;
  float e = __fmaf_rn(-xy, y0, 1.0f);
  return __fmaf_rn(0.5f * y0, e, y0);      // one Newton step
#else
  (void)x;
  return y0;
#endif
}
// 1/2 as a multiplier operand (constant bank on the device, see above)
JPD_HD double half_const(double) {
#if defined(__CUDA_ARCH__) && !defined(JPD_NO_CONST_BANK)
  return jpd_c_half;
#else
  return 0.5;
#endif
}
JPD_HD float half_const(float) { return 0.5f; }
JPD_HD double rsqrt_refined(double x) { return rsqrt_correct(x, rsqrt_seed(x)); }
JPD_HD float rsqrt_refined(float x) { return rsqrt_correct(x, rsqrt_seed(x)); }

// Is the off-diagonal `a` still worth rotating against its two diagonal entries?
// Reference test (jacobi_pd.hpp:191): skip iff  (mii + a == mii) && (mjj + a == mjj),
// i.e. |a| below half an ulp of BOTH diagonals.  Here the same decision is taken on
// the integer pipe from the exponent fields: |a| < 2^(Ea+1) <= 2^(E-kMant-1) = ulp/2
// whenever Ea + kMant + 2 <= E.  It never skips a pair the reference test would
// rotate; in the one-binade band where the two differ it rotates once more
// (harmless: the rotation is then ~identity).  a == 0 is always skipped
// (jacobi_pd.hpp:196), a zero diagonal never lets a nonzero `a` be skipped, and
// Inf/NaN in `a` count as "rotate" so that they poison the output visibly.
template <typename T>
JPD_HD bool pair_needs_rotation(T a, T mii, T mjj) {
  const int ea = expo_field(a), ei = expo_field(mii), ej = expo_field(mjj);
  const int emin = ei < ej ? ei : ej;
  const bool negligible = (ea + Fp<T>::kMantBits + 2 <= emin);
  return !is_zero(a) && !negligible;
}

// Rotation (c, s, t) = (cos, sin, tan) of the angle that zeroes the (i,j) entry.
// Same angle as the reference's CalcRot (|t| <= 1, the smaller root of
// t^2 + 2 kappa t - 1 = 0, kappa = d / 2a, d = mjj - mii; d == 0 gives t = +1),
// obtained through the double-angle identities instead of kappa:
//     x = d^2 + (2a)^2,  p = x^-1/2,  cos2θ = rho = |d| p,  sin2θ = sp = 2a p sgn(d)
//     v = (1 + rho)/2 = c^2,  u = v^-1/2,  c = v u,  s = sp u / 2,  t = s u
// => 2 refined rsqrt and ~10 FMA-class ops, no division, no cancellation (s comes
// from sin2θ / 2c, never from 1 - c^2).  d and a are pre-scaled by a power of two
// when x could overflow or underflow (warp-uniform rare path): c, s, t are
// scale-invariant, so huge / tiny / denormal inputs keep full accuracy.
// (u = 1/c is returned as well: the scaled-rotation tiers track reciprocal row scales with it)
template <typename T>
JPD_HD void rotation_from_scaled(T d, T a, bool d_is_zero, bool d_neg, T& c, T& s, T& t, T& u) {
  const T a2 = a + a;
  const T x = jfma(a2, a2, d * d);
  const T p = rsqrt_refined(x);
  const T rho = jabs(d) * p;
  T sp = a2 * p;
  // sign(t) = sign(d) * sign(a); the reference takes t = +1 when d == 0 whatever the
  // sign of a (jacobi_pd.hpp:237-239), keep that so n == 2 agrees even unsorted.
  const bool flip = d_is_zero ? sign_bit(a) : d_neg;
  sp = flip ? -sp : sp;
  const T v = jfma(T(0.5), rho, T(0.5));
  u = rsqrt_refined(v);
  c = v * u;
  s = (T(0.5) * sp) * u;
  t = s * u;
}
template <typename T>
JPD_HD void rotation_from_scaled(T d, T a, bool d_is_zero, bool d_neg, T& c, T& s, T& t) {
  T u;
  rotation_from_scaled(d, a, d_is_zero, d_neg, c, s, t, u);
}

template <typename T>
JPD_HD bool rotation_range_unsafe(T d, T a) {
  const int ed = expo_field(d), ea = expo_field(a);
  const int em = ed > ea ? ed : ea;
  return (em > Fp<T>::kBias + Fp<T>::kSafe) || (em < Fp<T>::kBias - Fp<T>::kSafe);
}

template <typename T>
JPD_HD void rescale_for_rotation(T& d, T& a) {
  const int ed = expo_field(d), ea = expo_field(a);
  int em = ed > ea ? ed : ea;
  if (em >= Fp<T>::kExpMask) return;  // Inf/NaN: let it propagate
  int field = 2 * Fp<T>::kBias - em;   // scale = 2^(bias - em): larger operand -> [1,2)
  if (field < 1) field = 1;
  if (field > Fp<T>::kExpMask - 1) field = Fp<T>::kExpMask - 1;
  const T sc = pow2_from_field(field, T(0));
  d *= sc;
  a *= sc;
}

// d = mjj - mii (computed by the caller, who also needs it for the range vote).
// `any_unsafe` must be uniform over the threads that execute this together (a warp
// vote on the device, the lane's own flag on the host) so the rare path costs
// nothing when no lane needs it.
template <typename T>
JPD_HD void jacobi_rotation(T d, T a, bool any_unsafe, T& c, T& s, T& t) {
  const bool d_is_zero = is_zero(d);
  const bool d_neg = sign_bit(d);
  if (any_unsafe) rescale_for_rotation(d, a);
  rotation_from_scaled(d, a, d_is_zero, d_neg, c, s, t);
}

// ---- shared by the parallel-ordering tiers (jpd_warp.cuh, jpd_cta.cuh, jpd_cluster.cuh) ----
// The "rotation" of those tiers also exchanges the two positions of its pair (odd-even
// ordering): applied to a 2-vector it is G = [[s, c], [c, -s]].

// Pivot pair with diagonal entries (mpp, mqq) and off-diagonal entry a: skip test
// (jacobi_pd.hpp:191), CalcRot (:233-256), new diagonal entries (:337-338) at the EXCHANGED
// positions.  Returns whether a rotation was needed; otherwise (c,s) = (1,0): a pure exchange.
template <typename T>
JPD_HD bool pivot_exchange_rotation(T mpp, T mqq, T a, T& c, T& s, T& npp, T& nqq) {
  c = T(1); s = T(0); npp = mqq; nqq = mpp;
  if (!pair_needs_rotation(a, mpp, mqq)) return false;
  const T d = mqq - mpp;
  T t;
  jacobi_rotation(d, a, rotation_range_unsafe(d, a), c, s, t);
  const T g = t * a;
  npp = mqq + g;
  nqq = mpp - g;
  return true;
}

// 2x2 block B = [[b00, b01], [b10, b11]] (rows: pair K, columns: pair L)  <-  G_K B G_L^T
// (ApplyRot :350-390 for both rotations at once): 8 multiplications + 8 FMAs.
template <typename T>
JPD_HD void exchange_rotate_block(T& b00, T& b01, T& b10, T& b11, T ck, T sk, T cl, T sl) {
  const T u00 = jfma(sk, b00, ck * b10), u01 = jfma(sk, b01, ck * b11);
  const T u10 = jfma(ck, b00, -(sk * b10)), u11 = jfma(ck, b01, -(sk * b11));
  b00 = jfma(sl, u00, cl * u01);
  b01 = jfma(cl, u00, -(sl * u01));
  b10 = jfma(sl, u10, cl * u11);
  b11 = jfma(cl, u10, -(sl * u11));
}

// 2-vector (x, y)  <-  G (x, y): rows of E (ApplyRotLeft :414-418), end strips of pattern B
template <typename T>
JPD_HD void exchange_rotate_pair(T& x, T& y, T c, T s) {
  const T nx = jfma(s, x, c * y);
  y = jfma(c, x, -(s * y));
  x = nx;
}

}  // namespace jpd
