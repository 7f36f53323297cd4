// jacobi_pd_cuda.hpp -- C++ front end of the B200 batched Jacobi eigensolver.
//
// (i)  jacobi_pd::Jacobi<Scalar,Vector,Matrix,ConstMatrix>: same public members and
//      semantics as the reference class (/root/reference/include/jacobi_pd.hpp:25-146:
//      ctor(n), SetSize, SortCriteria :56-62, Diagonalize :76-81, copy/move/swap/assign
//      :141-145), so existing callers -- LAMMPS' 3x3 inertia tensors, colvars' 4x4
//      quaternion overlap matrices, the reference's own tests/test_jacobi.cpp -- build
//      unchanged.  Diagonalize marshals ONE matrix through the host entry point of the C ABI
//      (jpd_diagonalize_batched_host_f64/_f32, include/jpd.h) and therefore runs on the GPU.
// (ii) jacobi_pd::BatchedJacobi<Scalar>: the new surface -- whole batches, device- or
//      host-resident, one or several GPUs -- thin inline wrappers over the same C ABI.
//
// Return value of Diagonalize as in the reference (:73-74, :214-219): iterations (> 0), or 0
// when max_num_sweeps was not enough and n > 1.  Differences, all documented in DESIGN.md:
// the pair order is a static round-robin instead of max-pivot (same eigen-decomposition
// within the reference's own accuracy; the iteration COUNT and the DO_NOT_SORT order differ
// for n > 2), and with calc_evecs == false `evec` is left untouched (the reference permutes
// whatever it contains, :504-506).  There is no CPU fallback: if libjpd.so cannot run the
// call (no device, CUDA error) a std::runtime_error is thrown.
#ifndef JPD_B200_JACOBI_PD_CUDA_HPP_
#define JPD_B200_JACOBI_PD_CUDA_HPP_

#include <cstddef>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "jpd.h"
#include "matrix_alloc_jpd.hpp"

namespace jacobi_pd {

namespace detail {

inline void check(int rc, const char *what) {
  if (rc == JPD_OK) return;
  std::string msg = std::string(what) + ": " + jpd_error_string(rc);
  const char *cuda = jpd_last_cuda_error();
  if (cuda && *cuda) msg += std::string(" [") + cuda + "]";
  throw std::runtime_error(msg);
}

template <typename Scalar> struct Abi;
template <> struct Abi<double> {
  static int device(const double *m, double *ev, double *vec, int *it, int n, long long b, int layout, int sort,
                    int calc, int sweeps, void *stream) {
    return jpd_diagonalize_batched_f64(m, ev, vec, it, n, b, layout, sort, calc, sweeps, stream);
  }
  static int host(const double *m, double *ev, double *vec, int *it, int n, long long b, int layout, int sort,
                  int calc, int sweeps, int device_id) {
    return jpd_diagonalize_batched_host_f64(m, ev, vec, it, n, b, layout, sort, calc, sweeps, device_id);
  }
  static int host_multi(const double *m, double *ev, double *vec, int *it, int n, long long b, int layout,
                        int sort, int calc, int sweeps, int n_dev, const int *devs) {
    return jpd_diagonalize_batched_host_multi_f64(m, ev, vec, it, n, b, layout, sort, calc, sweeps, n_dev, devs);
  }
  static int quaternion(const double *c, double *lam, double *q, int *it, long long b, int layout, int all,
                        int sweeps, void *stream) {
    return jpd_quaternion_from_correlation_f64(c, lam, q, it, b, layout, all, sweeps, stream);
  }
  static int axes(const double *i6, double *mom, double *ax, int *it, long long b, int layout, int sweeps,
                  void *stream) {
    return jpd_principal_axes_f64(i6, mom, ax, it, b, layout, sweeps, stream);
  }
  static int host_packed(const double *m, double *ev, double *vec, int *it, int n, long long b, int layout, int sort,
                         int calc, int sweeps, int device_id) {
    return jpd_diagonalize_batched_host_packed_f64(m, ev, vec, it, n, b, layout, sort, calc, sweeps, device_id);
  }
  static int quaternion_host(const double *c, double *lam, double *q, int *it, long long b, int layout, int all,
                             int sweeps, int device_id) {
    return jpd_quaternion_from_correlation_host_f64(c, lam, q, it, b, layout, all, sweeps, device_id);
  }
  static int axes_host(const double *i6, double *mom, double *ax, int *it, long long b, int layout, int sweeps,
                       int device_id) {
    return jpd_principal_axes_host_f64(i6, mom, ax, it, b, layout, sweeps, device_id);
  }
};
template <> struct Abi<float> {
  static int device(const float *m, float *ev, float *vec, int *it, int n, long long b, int layout, int sort,
                    int calc, int sweeps, void *stream) {
    return jpd_diagonalize_batched_f32(m, ev, vec, it, n, b, layout, sort, calc, sweeps, stream);
  }
  static int host(const float *m, float *ev, float *vec, int *it, int n, long long b, int layout, int sort,
                  int calc, int sweeps, int device_id) {
    return jpd_diagonalize_batched_host_f32(m, ev, vec, it, n, b, layout, sort, calc, sweeps, device_id);
  }
  static int host_multi(const float *m, float *ev, float *vec, int *it, int n, long long b, int layout,
                        int sort, int calc, int sweeps, int n_dev, const int *devs) {
    return jpd_diagonalize_batched_host_multi_f32(m, ev, vec, it, n, b, layout, sort, calc, sweeps, n_dev, devs);
  }
  static int quaternion(const float *c, float *lam, float *q, int *it, long long b, int layout, int all,
                        int sweeps, void *stream) {
    return jpd_quaternion_from_correlation_f32(c, lam, q, it, b, layout, all, sweeps, stream);
  }
  static int axes(const float *i6, float *mom, float *ax, int *it, long long b, int layout, int sweeps,
                  void *stream) {
    return jpd_principal_axes_f32(i6, mom, ax, it, b, layout, sweeps, stream);
  }
  static int host_packed(const float *m, float *ev, float *vec, int *it, int n, long long b, int layout, int sort,
                         int calc, int sweeps, int device_id) {
    return jpd_diagonalize_batched_host_packed_f32(m, ev, vec, it, n, b, layout, sort, calc, sweeps, device_id);
  }
  static int quaternion_host(const float *c, float *lam, float *q, int *it, long long b, int layout, int all,
                             int sweeps, int device_id) {
    return jpd_quaternion_from_correlation_host_f32(c, lam, q, it, b, layout, all, sweeps, device_id);
  }
  static int axes_host(const float *i6, float *mom, float *ax, int *it, long long b, int layout, int sweeps,
                       int device_id) {
    return jpd_principal_axes_host_f32(i6, mom, ax, it, b, layout, sweeps, device_id);
  }
};

}  // namespace detail

/// Memory layout of a batch (see include/jpd.h).
enum class Layout : int { AoS = JPD_LAYOUT_AOS, SoAInterleaved = JPD_LAYOUT_SOA };

// ---------------------------------------------------------------------------------------
// (i) drop-in scalar class
// ---------------------------------------------------------------------------------------
template <typename Scalar, typename Vector, typename Matrix, typename ConstMatrix = Matrix>
class Jacobi {
  // The reference is generic in Scalar (jacobi_pd.hpp:25); the device computes in fp64 or fp32
  // only.  Any other Scalar (long double, a multiprecision type, ...) would have to be rounded
  // to double behind the caller's back, so it is refused at compile time.
  static_assert(std::is_same<Scalar, double>::value || std::is_same<Scalar, float>::value,
                "the GPU path computes in fp64 or fp32");
  int n;                         // matrix size (rows)
  int device_id;                 // CUDA device the calls run on
  std::vector<Scalar> in_;       // contiguous n*n staging copy of the caller's matrix
  std::vector<Scalar> evals_;    // n
  std::vector<Scalar> evecs_;    // n*n, eigenvector k in row k

 public:
  /// Same enumerators and values as the reference (jacobi_pd.hpp:56-62).
  using SortCriteria = enum {
    DO_NOT_SORT = JPD_DO_NOT_SORT,
    SORT_DECREASING_EVALS = JPD_SORT_DECREASING_EVALS,
    SORT_INCREASING_EVALS = JPD_SORT_INCREASING_EVALS,
    SORT_DECREASING_ABS_EVALS = JPD_SORT_DECREASING_ABS_EVALS,
    SORT_INCREASING_ABS_EVALS = JPD_SORT_INCREASING_ABS_EVALS
  };

  explicit Jacobi(int n_rows = 0) : n(0), device_id(0) { SetSize(n_rows); }
  ~Jacobi() = default;

  /// Size of the matrices to diagonalise later (jacobi_pd.hpp:42); (re)allocates staging.
  void SetSize(int n_rows) {
    n = n_rows > 0 ? n_rows : 0;
    const std::size_t nn = static_cast<std::size_t>(n) * n;
    in_.assign(nn, Scalar(0));
    evals_.assign(static_cast<std::size_t>(n), Scalar(0));
    evecs_.assign(nn, Scalar(0));
  }

  /// Not in the reference: choose the CUDA device (default 0).
  void SetDevice(int device) { device_id = device; }

  /// Eigenvalues -> eval[0..n), eigenvectors -> ROWS of evec; only mat[i][j], j >= i is read
  /// (jacobi_pd.hpp:64-81, :167-171).  Returns iterations (> 0) or 0 if not converged.
  int Diagonalize(ConstMatrix mat, Vector eval, Matrix evec,
                  SortCriteria sort_criteria = SORT_DECREASING_EVALS, bool calc_evecs = true,
                  int max_num_sweeps = 50) {
    if (n <= 0) return 1;  // the reference's loop does not run either and it returns n_iters = 1
    for (int i = 0; i < n; ++i) {
      for (int j = i; j < n; ++j) {
        const Scalar v = mat[i][j];
        in_[static_cast<std::size_t>(i) * n + j] = v;
        in_[static_cast<std::size_t>(j) * n + i] = v;
      }
    }
    int iters = 0;
    detail::check(detail::Abi<Scalar>::host(in_.data(), evals_.data(), calc_evecs ? evecs_.data() : nullptr,
                                           &iters, n, 1, JPD_LAYOUT_AOS, static_cast<int>(sort_criteria),
                                           calc_evecs ? 1 : 0, max_num_sweeps, device_id),
                  "jacobi_pd::Jacobi::Diagonalize");
    for (int i = 0; i < n; ++i) eval[i] = evals_[static_cast<std::size_t>(i)];
    if (calc_evecs)
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) evec[i][j] = evecs_[static_cast<std::size_t>(i) * n + j];
    return iters;
  }

  // value semantics as the reference (jacobi_pd.hpp:141-145): deep copy, move, swap, copy-swap
  Jacobi(const Jacobi &source) = default;
  Jacobi(Jacobi &&other) noexcept : Jacobi() { this->swap(other); }
  void swap(Jacobi &other) noexcept {
    std::swap(n, other.n);
    std::swap(device_id, other.device_id);
    in_.swap(other.in_);
    evals_.swap(other.evals_);
    evecs_.swap(other.evecs_);
  }
  Jacobi &operator=(Jacobi source) {
    this->swap(source);
    return *this;
  }
};

// ---------------------------------------------------------------------------------------
// (ii) batched front end
// ---------------------------------------------------------------------------------------
/// Many independent n x n symmetric matrices laid out contiguously (Layout above).
/// Trailing arguments and per-matrix results mean exactly what they mean for
/// Jacobi::Diagonalize; n_iters[b] carries the return value of matrix b.
template <typename Scalar>
struct BatchedJacobi {
  static_assert(std::is_same<Scalar, double>::value || std::is_same<Scalar, float>::value,
                "the GPU path computes in fp64 or fp32");

  /// Device-resident batch on the current CUDA device; asynchronous on `cuda_stream`.
  static void Diagonalize(const Scalar *d_mat, Scalar *d_evals, Scalar *d_evecs, int *d_n_iters, int n,
                          long long batch, Layout layout = Layout::SoAInterleaved,
                          int sort_criteria = JPD_SORT_DECREASING_EVALS, bool calc_evecs = true,
                          int max_num_sweeps = 50, void *cuda_stream = nullptr) {
    detail::check(detail::Abi<Scalar>::device(d_mat, d_evals, d_evecs, d_n_iters, n, batch,
                                             static_cast<int>(layout), sort_criteria, calc_evecs ? 1 : 0,
                                             max_num_sweeps, cuda_stream),
                  "jacobi_pd::BatchedJacobi::Diagonalize");
  }

  /// Host-resident batch streamed through one GPU (H2D / kernel / D2H overlapped); synchronous.
  static void DiagonalizeHost(const Scalar *h_mat, Scalar *h_evals, Scalar *h_evecs, int *h_n_iters, int n,
                              long long batch, Layout layout = Layout::AoS,
                              int sort_criteria = JPD_SORT_DECREASING_EVALS, bool calc_evecs = true,
                              int max_num_sweeps = 50, int device = 0) {
    detail::check(detail::Abi<Scalar>::host(h_mat, h_evals, h_evecs, h_n_iters, n, batch,
                                           static_cast<int>(layout), sort_criteria, calc_evecs ? 1 : 0,
                                           max_num_sweeps, device),
                  "jacobi_pd::BatchedJacobi::DiagonalizeHost");
  }

  /// Host-resident batch sliced over the GPUs of one box: independent contiguous slices,
  /// one host thread and three streams per device, no inter-GPU traffic.
  static void DiagonalizeHostMultiGpu(const Scalar *h_mat, Scalar *h_evals, Scalar *h_evecs, int *h_n_iters,
                                      int n, long long batch, int n_devices, Layout layout = Layout::AoS,
                                      int sort_criteria = JPD_SORT_DECREASING_EVALS, bool calc_evecs = true,
                                      int max_num_sweeps = 50, const int *devices = nullptr) {
    detail::check(detail::Abi<Scalar>::host_multi(h_mat, h_evals, h_evecs, h_n_iters, n, batch,
                                                 static_cast<int>(layout), sort_criteria, calc_evecs ? 1 : 0,
                                                 max_num_sweeps, n_devices, devices),
                  "jacobi_pd::BatchedJacobi::DiagonalizeHostMultiGpu");
  }

  /// colvars optimal rotation, fused: 3x3 correlation matrices (row-major, 9 scalars each) ->
  /// largest eigenvalue of the 4x4 overlap matrix and its eigenvector, the rotation
  /// quaternion (q0 >= 0); with all_eigenpairs the 4 eigenvalues (decreasing) and 4x4
  /// eigenvector rows.  Device pointers, asynchronous on `cuda_stream` (include/jpd.h).
  static void QuaternionFromCorrelation(const Scalar *d_corr, Scalar *d_lambda, Scalar *d_quat, int *d_n_iters,
                                        long long batch, Layout layout = Layout::AoS,
                                        bool all_eigenpairs = false, int max_num_sweeps = 50,
                                        void *cuda_stream = nullptr) {
    detail::check(detail::Abi<Scalar>::quaternion(d_corr, d_lambda, d_quat, d_n_iters, batch,
                                                 static_cast<int>(layout), all_eigenpairs ? 1 : 0,
                                                 max_num_sweeps, cuda_stream),
                  "jacobi_pd::BatchedJacobi::QuaternionFromCorrelation");
  }

  /// Host-resident batch given as packed upper triangles (n(n+1)/2 scalars per matrix -- all the
  /// reference reads, jacobi_pd.hpp:167-171): less PCIe traffic than DiagonalizeHost.
  static void DiagonalizeHostPacked(const Scalar *h_upper, Scalar *h_evals, Scalar *h_evecs, int *h_n_iters, int n,
                                    long long batch, Layout layout = Layout::AoS,
                                    int sort_criteria = JPD_SORT_DECREASING_EVALS, bool calc_evecs = true,
                                    int max_num_sweeps = 50, int device = 0) {
    detail::check(detail::Abi<Scalar>::host_packed(h_upper, h_evals, h_evecs, h_n_iters, n, batch,
                                                  static_cast<int>(layout), sort_criteria, calc_evecs ? 1 : 0,
                                                  max_num_sweeps, device),
                  "jacobi_pd::BatchedJacobi::DiagonalizeHostPacked");
  }

  /// QuaternionFromCorrelation / PrincipalAxes with HOST buffers, streamed through one GPU like
  /// DiagonalizeHost (colvars: 72 B up + 44 B down per frame instead of 128 + 168 B).
  static void QuaternionFromCorrelationHost(const Scalar *h_corr, Scalar *h_lambda, Scalar *h_quat, int *h_n_iters,
                                            long long batch, Layout layout = Layout::AoS,
                                            bool all_eigenpairs = false, int max_num_sweeps = 50, int device = 0) {
    detail::check(detail::Abi<Scalar>::quaternion_host(h_corr, h_lambda, h_quat, h_n_iters, batch,
                                                      static_cast<int>(layout), all_eigenpairs ? 1 : 0,
                                                      max_num_sweeps, device),
                  "jacobi_pd::BatchedJacobi::QuaternionFromCorrelationHost");
  }
  static void PrincipalAxesHost(const Scalar *h_inertia6, Scalar *h_moments, Scalar *h_axes, int *h_n_iters,
                                long long batch, Layout layout = Layout::AoS, int max_num_sweeps = 50,
                                int device = 0) {
    detail::check(detail::Abi<Scalar>::axes_host(h_inertia6, h_moments, h_axes, h_n_iters, batch,
                                                static_cast<int>(layout), max_num_sweeps, device),
                  "jacobi_pd::BatchedJacobi::PrincipalAxesHost");
  }

  /// Page-lock / release a caller-owned host buffer so that the host entries overlap copies and kernels.
  static void HostRegister(void *h_ptr, std::size_t bytes) {
    detail::check(jpd_host_register(h_ptr, static_cast<unsigned long long>(bytes)), "jacobi_pd::BatchedJacobi::HostRegister");
  }
  static void HostUnregister(void *h_ptr) {
    detail::check(jpd_host_unregister(h_ptr), "jacobi_pd::BatchedJacobi::HostUnregister");
  }

  /// LAMMPS rigid bodies, fused: (Ixx, Iyy, Izz, Iyz, Ixz, Ixy) per body -> principal moments
  /// (decreasing) and a right-handed frame with the principal axes as COLUMNS.
  static void PrincipalAxes(const Scalar *d_inertia6, Scalar *d_moments, Scalar *d_axes, int *d_n_iters,
                            long long batch, Layout layout = Layout::AoS, int max_num_sweeps = 50,
                            void *cuda_stream = nullptr) {
    detail::check(detail::Abi<Scalar>::axes(d_inertia6, d_moments, d_axes, d_n_iters, batch,
                                           static_cast<int>(layout), max_num_sweeps, cuda_stream),
                  "jacobi_pd::BatchedJacobi::PrincipalAxes");
  }
};

}  // namespace jacobi_pd

#endif  // JPD_B200_JACOBI_PD_CUDA_HPP_
