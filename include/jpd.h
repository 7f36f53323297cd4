/* jpd.h -- C ABI of libjpd.so, the B200 (sm_100a) batched Jacobi eigensolver.
 *
 * Drop-in boundary for ONE path of jewettaij/jacobi_pd:
 *
 *   int Jacobi<Scalar,Vector,Matrix,ConstMatrix>::Diagonalize(
 *         ConstMatrix mat, Vector eval, Matrix evec,
 *         SortCriteria sort_criteria = SORT_DECREASING_EVALS,
 *         bool calc_evecs = true, int max_num_sweeps = 50);
 *                                   (reference: include/jacobi_pd.hpp:76-81, impl :159-220)
 *
 * The reference is a header-only C++ template, it has no FFI of its own; these entry
 * points are what a binding for that member function would call.  Plain pointers and
 * sizes only, no C++ or torch types.  INTEGRATION.md shows the reference-side stub.
 *
 * Semantics kept from the reference (per matrix):
 *   - only mat[i][j], j >= i is read                         (jacobi_pd.hpp:167-171)
 *   - eigenvector k is ROW k of evecs                        (jacobi_pd.hpp:69)
 *   - sort_criteria values 0..4 = DO_NOT_SORT, SORT_DECREASING_EVALS,
 *     SORT_INCREASING_EVALS, SORT_DECREASING_ABS_EVALS, SORT_INCREASING_ABS_EVALS;
 *     ties resolve as the reference's selection sort does   (jacobi_pd.hpp:56-62, 477-508)
 *   - n_iters[b] > 0  : converged (rotations performed + 1)
 *     n_iters[b] == 0 : max_num_sweeps exhausted and n > 1   (jacobi_pd.hpp:214-219)
 * Deliberate deviations (DESIGN.md "deviations"): pair order is static (round-robin / odd-even
 * parallel ordering), not max-pivot, so rotation counts and the DO_NOT_SORT order differ for n > 2; when
 * calc_evecs == 0 the evecs buffer may be NULL and is never touched.
 *
 * Layouts (argument `layout`):
 *   JPD_LAYOUT_AOS (0)  mat[b*n*n + i*n + j]   evals[b*n + k]   evecs[b*n*n + k*n + v]
 *   JPD_LAYOUT_SOA (1)  mat[(i*n+j)*batch + b] evals[k*batch+b] evecs[(k*n+v)*batch + b]
 *                       (interleaved structure of arrays: coalesced for thread-per-matrix;
 *                        lower-triangle planes of `mat` are never read)
 *
 * Return value of every function: 0 on success, negative JPD_ERR_* otherwise.  Nothing
 * throws.  There is NO CPU fallback: without a CUDA device the calls fail with
 * JPD_ERR_CUDA.
 */
#ifndef JPD_H_
#define JPD_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JPD_LAYOUT_AOS 0
#define JPD_LAYOUT_SOA 1
/* adapter-only layouts (jpd_convert_layout_*): upper triangle packed row-major,
 * t(i,j) = i*n - i(i-1)/2 + (j-i), T = n(n+1)/2 scalars per matrix; and an array of
 * device pointers to scattered row-major n x n matrices (source only) */
#define JPD_LAYOUT_PACKED_AOS 2   /* m[b*T + t(i,j)]     */
#define JPD_LAYOUT_PACKED_SOA 3   /* m[t(i,j)*batch + b] */
#define JPD_LAYOUT_POINTERS 4     /* ((const S* const*)src)[b][i*n + j] */

#define JPD_DO_NOT_SORT 0
#define JPD_SORT_DECREASING_EVALS 1
#define JPD_SORT_INCREASING_EVALS 2
#define JPD_SORT_DECREASING_ABS_EVALS 3
#define JPD_SORT_INCREASING_ABS_EVALS 4

#define JPD_OK 0
#define JPD_ERR_INVALID_ARG (-1)
#define JPD_ERR_UNSUPPORTED (-2)
#define JPD_ERR_CUDA (-3)
#define JPD_ERR_ALLOC (-4)

#define JPD_MAX_N 1024

#if defined(__GNUC__)
#define JPD_API __attribute__((visibility("default")))
#else
#define JPD_API
#endif

/* ---- device-resident batches (the hot path) --------------------------------------
 * All pointers are DEVICE pointers on the current device; the call is asynchronous on
 * `cuda_stream` (a cudaStream_t, NULL = default stream).  d_evecs may be NULL iff
 * calc_evecs == 0; d_n_iters may be NULL.  Replaces a loop of Diagonalize() calls
 * (jacobi_pd.hpp:159-220) over `batch` matrices of size n x n. */
JPD_API int jpd_diagonalize_batched_f64(const double* d_mat, double* d_evals, double* d_evecs,
                                int* d_n_iters, int n, long long batch, int layout,
                                int sort_criteria, int calc_evecs, int max_num_sweeps,
                                void* cuda_stream);
JPD_API int jpd_diagonalize_batched_f32(const float* d_mat, float* d_evals, float* d_evecs,
                                int* d_n_iters, int n, long long batch, int layout,
                                int sort_criteria, int calc_evecs, int max_num_sweeps,
                                void* cuda_stream);

/* ---- host-resident batches ---------------------------------------------------------
 * Same contract with HOST pointers (pinned memory overlaps best, pageable works; see
 * jpd_host_register).  The batch is streamed through device `device` in chunks: H2D copy,
 * kernel and D2H copy of consecutive chunks overlap on three streams.  Synchronous: results
 * are in the host buffers on return, also after an error (nothing stays in flight).  The
 * caller's current CUDA device is restored on return.  This is what the drop-in C++ class Jacobi<>::Diagonalize
 * (include/jacobi_pd_cuda.hpp) calls with batch == 1. */
JPD_API int jpd_diagonalize_batched_host_f64(const double* h_mat, double* h_evals, double* h_evecs,
                                     int* h_n_iters, int n, long long batch, int layout,
                                     int sort_criteria, int calc_evecs, int max_num_sweeps,
                                     int device);
JPD_API int jpd_diagonalize_batched_host_f32(const float* h_mat, float* h_evals, float* h_evecs,
                                     int* h_n_iters, int n, long long batch, int layout,
                                     int sort_criteria, int calc_evecs, int max_num_sweeps,
                                     int device);

/* Same with the packed upper triangle as input (what the reference actually reads,
 * jacobi_pd.hpp:167-171): n(n+1)/2 scalars per matrix over PCIe instead of n*n.
 * layout 0: h_upper[b*T + t(i,j)], layout 1: h_upper[t(i,j)*batch + b], with
 * t(i,j) = i*n - i(i-1)/2 + (j-i), j >= i, T = n(n+1)/2; outputs as above. */
JPD_API int jpd_diagonalize_batched_host_packed_f64(const double* h_upper, double* h_evals, double* h_evecs,
                                            int* h_n_iters, int n, long long batch, int layout,
                                            int sort_criteria, int calc_evecs, int max_num_sweeps,
                                            int device);
JPD_API int jpd_diagonalize_batched_host_packed_f32(const float* h_upper, float* h_evals, float* h_evecs,
                                            int* h_n_iters, int n, long long batch, int layout,
                                            int sort_criteria, int calc_evecs, int max_num_sweeps,
                                            int device);
/* Page-locks a caller-owned buffer (cudaHostRegister, portable) so that the host entries
 * overlap their copies with the kernels; pageable memory works but serialises them.
 * Registering costs ~0.3 ms per MB once -- worth it for buffers that are reused. */
JPD_API int jpd_host_register(void* h_ptr, unsigned long long bytes);
JPD_API int jpd_host_unregister(void* h_ptr);

/* ---- host-resident batches over several GPUs of one box ----------------------------
 * Slices [0,batch) into n_devices contiguous parts (part g = [g*ceil(batch/G), ...)) and
 * runs each part on devices[g] (NULL => 0..n_devices-1) from its own host thread.  The
 * matrices are independent, so there is no inter-GPU traffic and no collective. */
JPD_API int jpd_diagonalize_batched_host_multi_f64(const double* h_mat, double* h_evals,
                                           double* h_evecs, int* h_n_iters, int n,
                                           long long batch, int layout, int sort_criteria,
                                           int calc_evecs, int max_num_sweeps,
                                           int n_devices, const int* devices);
JPD_API int jpd_diagonalize_batched_host_multi_f32(const float* h_mat, float* h_evals, float* h_evecs,
                                           int* h_n_iters, int n, long long batch, int layout,
                                           int sort_criteria, int calc_evecs,
                                           int max_num_sweeps, int n_devices,
                                           const int* devices);

/* ---- fused producers for the two callers the reference names (README.md:124-128) -----
 * The step before Diagonalize, done in registers by the thread-per-matrix kernel, so that
 * the 3x3 / 4x4 matrix never exists in memory.  Device pointers,
 * asynchronous on `cuda_stream`, same return codes and n_iters semantics as above.
 *
 * colvars optimal rotation: d_corr holds 3x3 correlation matrices C = sum_k x_k y_k^T,
 * row-major (xx xy xz yx yy yz zx zy zz): AoS corr[b*9 + e] or SoA corr[e*batch + b].  The
 * 4x4 overlap matrix S (Horn 1987; the formula of jpd_fill_synthetic kind 1, include/jpd_bench.h) is
 * diagonalised as Diagonalize(S, ..., SORT_DECREASING_EVALS) would (jacobi_pd.hpp:159-220).
 *   all_eigenpairs == 0: d_lambda[b] = largest eigenvalue, d_quat = its eigenvector with
 *                        q0 >= 0 (AoS quat[b*4 + v], SoA quat[v*batch + b])
 *   all_eigenpairs != 0: d_lambda / d_quat = evals (4) / evecs (4x4, rows) exactly as
 *                        jpd_diagonalize_batched would lay them out for n = 4 */
JPD_API int jpd_quaternion_from_correlation_f64(const double* d_corr, double* d_lambda, double* d_quat,
                                        int* d_n_iters, long long batch, int layout,
                                        int all_eigenpairs, int max_num_sweeps, void* cuda_stream);
JPD_API int jpd_quaternion_from_correlation_f32(const float* d_corr, float* d_lambda, float* d_quat,
                                        int* d_n_iters, long long batch, int layout,
                                        int all_eigenpairs, int max_num_sweeps, void* cuda_stream);
/* the same with HOST buffers, streamed through `device` like jpd_diagonalize_batched_host_*:
 * 72 B up + 44 B down per matrix (fp64, leading eigenpair) instead of 128 + 168 B */
JPD_API int jpd_quaternion_from_correlation_host_f64(const double* h_corr, double* h_lambda, double* h_quat,
                                             int* h_n_iters, long long batch, int layout,
                                             int all_eigenpairs, int max_num_sweeps, int device);
JPD_API int jpd_quaternion_from_correlation_host_f32(const float* h_corr, float* h_lambda, float* h_quat,
                                             int* h_n_iters, long long batch, int layout,
                                             int all_eigenpairs, int max_num_sweeps, int device);
/* LAMMPS rigid bodies: d_inertia6 holds (Ixx, Iyy, Izz, Iyz, Ixz, Ixy) per body (AoS
 * in[b*6 + e] or SoA in[e*batch + b]).  Diagonalize(n = 3, SORT_DECREASING_EVALS), then the
 * post-step of the caller: d_moments (3, decreasing), d_axes[i*3 + k] = component i of
 * principal axis k (eigenvectors as COLUMNS), third axis flipped if needed so that
 * (ex x ey) . ez > 0.  AoS axes[b*9 + i*3 + k], SoA axes[(i*3+k)*batch + b]. */
JPD_API int jpd_principal_axes_f64(const double* d_inertia6, double* d_moments, double* d_axes,
                           int* d_n_iters, long long batch, int layout, int max_num_sweeps,
                           void* cuda_stream);
JPD_API int jpd_principal_axes_f32(const float* d_inertia6, float* d_moments, float* d_axes,
                           int* d_n_iters, long long batch, int layout, int max_num_sweeps,
                           void* cuda_stream);

JPD_API int jpd_principal_axes_host_f64(const double* h_inertia6, double* h_moments, double* h_axes,
                                int* h_n_iters, long long batch, int layout, int max_num_sweeps,
                                int device);
JPD_API int jpd_principal_axes_host_f32(const float* h_inertia6, float* h_moments, float* h_axes,
                                int* h_n_iters, long long batch, int layout, int max_num_sweeps,
                                int device);

/* ---- layout adapters ---------------------------------------------------
 * Converts a device-resident batch of n x n matrices between the layouts above with one
 * tiled-transpose kernel (32 consecutive scalars per warp on both sides).  Sources: all five
 * layouts; targets: 0..3.  A packed source is mirrored into the lower triangle of a full
 * target; a packed target receives the upper triangle (as the reference reads only
 * mat[i][j], j >= i, jacobi_pd.hpp:167-171).  Also transposes eigenvector outputs between
 * AoS and SoA.  Callers holding packed or scattered matrices run it once in front of
 * jpd_diagonalize_batched_*. */
JPD_API int jpd_convert_layout_f64(const void* d_src, int src_layout, double* d_dst, int dst_layout,
                           int n, long long batch, void* cuda_stream);
JPD_API int jpd_convert_layout_f32(const void* d_src, int src_layout, float* d_dst, int dst_layout,
                           int n, long long batch, void* cuda_stream);

/* slice g of G of a batch: first index and length (the partition used above) */
JPD_API void jpd_slice(long long batch, int n_parts, int part, long long* first, long long* count);

/* ---- bookkeeping -------------------------------------------------------------------- */
/* algorithmic bytes per matrix: elem*(n(n+1)/2 + n + (calc_evecs? n*n : 0)) + 4   */
JPD_API long long jpd_algorithmic_bytes(int n, int elem_size, int calc_evecs);
/* which kernel tier the dispatcher picks: 1 thread/matrix (n <= 6), 2 warp/matrix
 * (n <= 32), 3 block/matrix with E in registers and M in shared memory (n <= 128),
 * 4 four-CTA cluster/matrix with M in distributed shared memory (n <= 256), 5 block/matrix
 * with a global (L2-resident) workspace (n <= JPD_MAX_N); negative on error */
JPD_API int jpd_kernel_tier(int n, int elem_size, int calc_evecs);
/* number of kernels launched by this library in this process so far */
JPD_API long long jpd_launch_count(void);
/* free cached device workspaces / staging buffers of all devices */
JPD_API int jpd_release(void);
JPD_API const char* jpd_error_string(int code);
/* text of the last CUDA error seen by this thread's calls ("" if none) */
JPD_API const char* jpd_last_cuda_error(void);
JPD_API int jpd_version(void);

#ifdef __cplusplus
}
#endif
#endif /* JPD_H_ */
