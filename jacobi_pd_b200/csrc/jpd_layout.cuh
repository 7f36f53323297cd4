// jpd_layout.cuh -- layout adapters: the on-device "wire formats" either
// side of the Diagonalize path (/root/reference/include/jacobi_pd.hpp:159-220 takes any
// [i][j]-indexable matrix, :167-171 reads only j >= i).  One tiled-transpose kernel converts
// a batch between
//   0 AoS              m[b*n*n + i*n + j]           (what the reference's callers hold)
//   1 SoA-interleaved  m[(i*n+j)*batch + b]         (native for thread-per-matrix)
//   2 packed AoS       m[b*T + t(i,j)], j >= i      (upper triangle only, T = n(n+1)/2)
//   3 packed SoA       m[t(i,j)*batch + b], j >= i
//   4 pointer array    ptrs[b][i*n + j]             (scattered matrices, source only)
// reading and writing 32 consecutive scalars per warp on both sides (32 x 33 shared tile).
// Packed sources are mirrored into the lower triangle of full targets.
// The callers' own sizes (n <= 16: nine or sixteen scalars per matrix) would leave most lanes of
// a 32-wide tile row idle on the AoS side, so AoS <-> SoA of small matrices goes through
// jpd_convert_flat_kernel instead: the AoS side is treated as ONE flat array (a warp always
// touches 32 consecutive scalars, across matrix boundaries), the SoA side as planes.
// Measured GB/s against the HBM copy peak: bench.py "layout_adapter".
#pragma once
#include "jpd_common.cuh"

namespace jpd {

enum : int { kLayoutAoS = 0, kLayoutSoA = 1, kLayoutPackedAoS = 2, kLayoutPackedSoA = 3, kLayoutPointers = 4 };

JPD_HD bool layout_batch_fast(int layout) { return layout == kLayoutSoA || layout == kLayoutPackedSoA; }
JPD_HD bool layout_packed(int layout) { return layout == kLayoutPackedAoS || layout == kLayoutPackedSoA; }

#if defined(__CUDACC__)

// slot of full element e = i*n + j inside one matrix of `layout`, or -1 (packed target, j < i)
__device__ __forceinline__ long long layout_slot(int layout, int n, int e, bool mirror) {
  if (!layout_packed(layout)) return e;
  int i = e / n, j = e - i * n;
  if (j < i) {
    if (!mirror) return -1;
    const int t = i; i = j; j = t;
  }
  return (long long)i * n - (long long)i * (i - 1) / 2 + (j - i);
}

template <typename T>
__global__ void __launch_bounds__(256)
jpd_convert_kernel(const void* __restrict__ src, int src_layout, T* __restrict__ dst, int dst_layout,
                   int n, long long batch) {
  __shared__ T tile[32][33];   // [matrix within tile][element within tile]
  const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
  const long long b0 = (long long)blockIdx.x * 32;
  const int e0 = blockIdx.y * 32;
  const int nn = n * n;
  const long long src_per = layout_packed(src_layout) ? (long long)n * (n + 1) / 2 : nn;
  const long long dst_per = layout_packed(dst_layout) ? (long long)n * (n + 1) / 2 : nn;
  // ---- load: fast index = tx
  for (int r = ty; r < 32; r += 8) {
    const bool bf = layout_batch_fast(src_layout);
    const long long b = b0 + (bf ? tx : r);
    const int e = e0 + (bf ? r : tx);
    T v = T(0);
    if (b < batch && e < nn) {
      const long long s = layout_slot(src_layout, n, e, true);
      if (src_layout == kLayoutPointers) v = static_cast<const T* const*>(src)[b][e];
      else if (bf) v = static_cast<const T*>(src)[s * batch + b];
      else v = static_cast<const T*>(src)[b * src_per + s];
    }
    if (bf) tile[tx][r] = v; else tile[r][tx] = v;
  }
  __syncthreads();
  // ---- store: fast index = tx
  for (int r = ty; r < 32; r += 8) {
    const bool bf = layout_batch_fast(dst_layout);
    const long long b = b0 + (bf ? tx : r);
    const int e = e0 + (bf ? r : tx);
    if (b < batch && e < nn) {
      const long long s = layout_slot(dst_layout, n, e, false);
      if (s >= 0) {
        const T v = bf ? tile[tx][r] : tile[r][tx];
        if (bf) dst[s * batch + b] = v; else dst[b * dst_per + s] = v;
      }
    }
  }
}

// AoS <-> SoA for small matrices (nn = n*n <= 256).  One block moves `tb` consecutive matrices
// (tb a power of two >= 32) through a shared tile with an odd row pitch (conflict-free on the
// plane side, where the 32 lanes of a warp read 32 different matrices).
template <typename T, bool TO_SOA>
__global__ void __launch_bounds__(256)
jpd_convert_flat_kernel(const T* __restrict__ src, T* __restrict__ dst, int nn, long long batch, int tb, int tb_shift) {
  extern __shared__ __align__(16) unsigned char jpd_flat_raw[];
  T* tile = reinterpret_cast<T*>(jpd_flat_raw);
  const int pitch = nn | 1;
  const long long b0 = (long long)blockIdx.x * tb;
  const long long left = batch - b0;
  const int nb = left < tb ? (int)left : tb;
  const int total = nb * nn;
  const int tid = threadIdx.x;
  const T* aos_in = src + b0 * nn;
  T* aos_out = dst + b0 * nn;
  // flat AoS index f = bl * nn + e, advanced by 256 per iteration without a division
  int e = tid % nn, bl = tid / nn;
  const int de = 256 % nn, db = 256 / nn;
  if (TO_SOA) {
    for (int f = tid; f < total; f += 256) {
      tile[bl * pitch + e] = aos_in[f];
      e += de; bl += db;
      if (e >= nn) { e -= nn; ++bl; }
    }
    __syncthreads();
    for (int idx = tid; idx < (nn << tb_shift); idx += 256) {
      const int pe = idx >> tb_shift, pb = idx & (tb - 1);
      if (pb < nb) dst[(long long)pe * batch + b0 + pb] = tile[pb * pitch + pe];
    }
  } else {
    for (int idx = tid; idx < (nn << tb_shift); idx += 256) {
      const int pe = idx >> tb_shift, pb = idx & (tb - 1);
      if (pb < nb) tile[pb * pitch + pe] = src[(long long)pe * batch + b0 + pb];
    }
    __syncthreads();
    for (int f = tid; f < total; f += 256) {
      aos_out[f] = tile[bl * pitch + e];
      e += de; bl += db;
      if (e >= nn) { e -= nn; ++bl; }
    }
  }
}

#endif  // __CUDACC__

}  // namespace jpd
