// Circular correlation / convolution in shared memory (HolE).
//
//   ccorr(a, b)_k = sum_i a_i b_{(i+k) mod d}          (skge/util.py:30-50)
//   cconv(a, b)_k = sum_i a_i b_{(k-i) mod d}          (skge/util.py:8-27)
//
// The reference goes through numpy's complex FFT; on the device both are
// evaluated as sliding dot products against a DOUBLED copy of the second
// operand held in shared memory (so no modulo in the inner loop):
//   ccorr(a, b)_k = sum_i a_i B2[i + k]                 B2 = [b, b]
//   cconv(a, b)_k = sum_i a_i Brev2[i + (d-k) mod d]    Brev2 = [b', b'], b'_m = b_{(d-m) mod d}
// Thread k of the CTA owns output k.
#pragma once
#include "common.cuh"

namespace skge {

// dst[0..2d) <- [src, src]
__device__ __forceinline__ void smem_load_doubled(float *dst, const float *__restrict__ src, int d) {
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    float v = __ldg(src + i);
    dst[i] = v;
    dst[i + d] = v;
  }
}
// dst[0..2d) <- [rev(src), rev(src)] with rev(src)_m = src_{(d-m) mod d}
__device__ __forceinline__ void smem_load_rev_doubled(float *dst, const float *__restrict__ src, int d) {
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    float v = __ldg(src + i);
    int m = (d - i) % d;
    dst[m] = v;
    dst[m + d] = v;
  }
}
__device__ __forceinline__ void smem_load(float *dst, const float *__restrict__ src, int d) {
  for (int i = threadIdx.x; i < d; i += blockDim.x) dst[i] = __ldg(src + i);
}

// sum_i a[i] * x2[i + off]   (a: d floats in smem, x2: 2d floats in smem)
__device__ __forceinline__ float sliding_dot(const float *a, const float *x2, int off, int d) {
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
  const float *x = x2 + off;
  int i = 0;
  for (; i + 4 <= d; i += 4) {
    acc0 = fmaf(a[i], x[i], acc0);
    acc1 = fmaf(a[i + 1], x[i + 1], acc1);
    acc2 = fmaf(a[i + 2], x[i + 2], acc2);
    acc3 = fmaf(a[i + 3], x[i + 3], acc3);
  }
  for (; i < d; ++i) acc0 = fmaf(a[i], x[i], acc0);
  return (acc0 + acc1) + (acc2 + acc3);
}

// Block-wide sum; `red` is >= 33 floats of shared memory. All threads get the result.
__device__ __forceinline__ float block_sum(float v, float *red) {
  v = warp_sum(v);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    float t = l < nw ? red[l] : 0.f;
    t = warp_sum(t);
    if (l == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

}  // namespace skge
