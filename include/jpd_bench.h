/* jpd_bench.h -- bench / test scaffolding exported by libjpd.so.  NOT part of the product ABI
 * (include/jpd.h): nothing here is needed to use the solver. */
#ifndef JPD_BENCH_H_
#define JPD_BENCH_H_
#include "jpd.h"
#ifdef __cplusplus
extern "C" {
#endif

/* ---- synthetic inputs (bench / tests; host twin: oracle/jpd_oracle.c) ---------------
 * Counter-based generator, identical bits on host and device:
 *   kind 0: upper-triangle entries iid U(-1,1), mirrored
 *   kind 1: n == 4, quaternion-superposition overlap matrix of a 3x3 U(-1,1) correlation
 *   kind 2: n == 3, that 3x3 correlation matrix itself (row-major, NOT symmetric; device only):
 *           the input of jpd_quaternion_from_correlation_* whose overlap matrix is kind 1
 * Writes matrices b0 .. b0+batch-1 of the stream at local indices 0..batch-1 (device ptr). */
JPD_API int jpd_fill_synthetic_f64(double* d_mat, int n, long long batch, int layout, int kind,
                           unsigned long long seed, long long b0, void* cuda_stream);
JPD_API int jpd_fill_synthetic_f32(float* d_mat, int n, long long batch, int layout, int kind,
                           unsigned long long seed, long long b0, void* cuda_stream);

/* ---- FMA-pipe probe (bench only) -------------------------------------------------------
 * Launches a register-only stream of independent FMAs (elem_size 8 = DFMA, 4 = FFMA) and
 * returns how many FMAs it issues (negative on error); bench.py times it with CUDA events
 * to obtain the measured FP64/FP32 peak used as the FP-bound roofline denominator. */
JPD_API long long jpd_fma_probe(int elem_size, int blocks, int threads, int iters, void* d_scratch,
                                void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* JPD_BENCH_H_ */
