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

// ---- register-blocked form --------------------------------------------------------------
// A thread owns FOUR consecutive outputs k0..k0+3 (k0 % 4 == 0) and walks the inputs four at a
// time: one 128-bit load of a[i0..i0+3] (a broadcast) and two of the window x2[i0+k0 .. +7]
// feed 16 FMAs, instead of two 32-bit loads per FMA.  Layout (d4 = d rounded up to 4):
//   a  : d4 floats, a[i] = 0 for i >= d
//   x2 : 2*d4 floats, x2[j] = x[j mod d]  (periodic, so i + k never needs a modulo)
// and cconv(a, b) = ccorr(rev(a), b) with rev(a)_m = a_{(d-m) mod d}, so one doubled copy of b
// serves both.  Outputs k >= d are garbage and must be discarded by the caller.
__host__ __device__ __forceinline__ int round4(int d) { return (d + 3) & ~3; }

__device__ __forceinline__ void smem_load_padded(float *dst, const float *__restrict__ src, int d, int d4) {
  for (int i = threadIdx.x; i < d4; i += blockDim.x) dst[i] = i < d ? __ldg(src + i) : 0.f;
}
__device__ __forceinline__ void smem_load_rev_padded(float *dst, const float *__restrict__ src, int d, int d4) {
  for (int i = threadIdx.x; i < d4; i += blockDim.x) dst[i] = i < d ? __ldg(src + (i == 0 ? 0 : d - i)) : 0.f;
}
__device__ __forceinline__ void smem_load_periodic(float *dst, const float *__restrict__ src, int d, int d4) {
  for (int j = threadIdx.x; j < 2 * d4; j += blockDim.x) dst[j] = __ldg(src + j % d);
}

// out[m] += sum_{i in [i_beg, i_end)} a[i] * x2[i + k0 + m], m = 0..3 (i_beg, i_end, k0 multiples of 4)
__device__ __forceinline__ void sliding_dot4(const float *a, const float *x2, int k0, int i_beg, int i_end,
                                             float (&out)[4]) {
  const float *x = x2 + k0;
  for (int i0 = i_beg; i0 < i_end; i0 += 4) {
    const float4 av = *reinterpret_cast<const float4 *>(a + i0);
    const float4 w0 = *reinterpret_cast<const float4 *>(x + i0);
    const float4 w1 = *reinterpret_cast<const float4 *>(x + i0 + 4);
    out[0] = fmaf(av.x, w0.x, fmaf(av.y, w0.y, fmaf(av.z, w0.z, fmaf(av.w, w0.w, out[0]))));
    out[1] = fmaf(av.x, w0.y, fmaf(av.y, w0.z, fmaf(av.z, w0.w, fmaf(av.w, w1.x, out[1]))));
    out[2] = fmaf(av.x, w0.z, fmaf(av.y, w0.w, fmaf(av.z, w1.x, fmaf(av.w, w1.y, out[2]))));
    out[3] = fmaf(av.x, w0.w, fmaf(av.y, w1.x, fmaf(av.z, w1.y, fmaf(av.w, w1.z, out[3]))));
  }
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
