// Shared-memory FFT building blocks for HolE (power-of-two d).
//
// Spectra of real rows are kept in a packed "half-complex" layout of exactly d floats:
//   float2 slot 0      = (X_0, X_{d/2})            (both real)
//   float2 slot f      = (Re X_f, Im X_f)          0 < f < d/2
// with X_f = sum_n x_n exp(-2 pi i f n / d).  Products of spectra are slot-wise (slot 0 holds
// two independent real products), sums of spectra are plain float sums, and
//   sum_k x_k y_k = (1/d) [X_0 Y_0 + X_h Y_h + 2 sum_{0<f<h} Re(X_f conj Y_f)]      (Parseval).
#pragma once
#include "common.cuh"

namespace skge {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {  // conj(a) * b
  return make_float2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);
}

// tw[m] = exp(-2 pi i m / N), m < N/2; index m < N by symmetry
__device__ __forceinline__ float2 tw_at(const float2 *tw, int m, int half) {
  float2 w = tw[m & (half - 1)];
  return m >= half ? make_float2(-w.x, -w.y) : w;
}

__device__ __forceinline__ void fill_twiddles(float2 *tw, int N, int tid, int nthreads) {
  for (int m = tid; m < N / 2; m += nthreads) {
    float sn, cs;
    sincospif(-2.0f * (float)m / (float)N, &sn, &cs);
    tw[m] = make_float2(cs, sn);
  }
}

// One complex transform of length n = 1 << logn by ONE WARP: Stockham autosort, radix-4
// stages (plus a leading radix-2 stage when logn is odd), ping-ponging between `in` and
// `out` (n float2 each, shared memory private to the warp).  The twiddle table belongs to a
// transform of length n * tw_stride (tw[m] = exp(-2 pi i m / (n * tw_stride))).  Returns the
// buffer that holds the result.  Callers must __syncwarp() after filling `in`.
template <bool INVERSE>
__device__ __forceinline__ float2 *warp_fft(float2 *in, float2 *out, const float2 *tw, int tw_stride, int logn,
                                            int lane) {
  const int N = 1 << logn, H = N >> 1, Qn = N >> 2;
  const int thalf = H * tw_stride;  // table entries
  int Ns = 1;
  if (logn & 1) {
    for (int j = lane; j < H; j += 32) {
      const float2 u0 = in[j], u1 = in[j + H];
      out[2 * j] = make_float2(u0.x + u1.x, u0.y + u1.y);
      out[2 * j + 1] = make_float2(u0.x - u1.x, u0.y - u1.y);
    }
    __syncwarp();
    float2 *t = in; in = out; out = t;
    Ns = 2;
  }
  for (; Ns < N; Ns <<= 2) {
    const int tstep = (N / (4 * Ns)) * tw_stride;
    for (int j = lane; j < Qn; j += 32) {
      const int k = j & (Ns - 1);
      const int j0 = ((j - k) << 2) + k;
      float2 w1 = tw_at(tw, k * tstep, thalf), w2 = tw_at(tw, 2 * k * tstep, thalf), w3 = tw_at(tw, 3 * k * tstep, thalf);
      if (INVERSE) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; }
      const float2 v0 = in[j];
      const float2 v1 = cmul(in[j + Qn], w1);
      const float2 v2 = cmul(in[j + 2 * Qn], w2);
      const float2 v3 = cmul(in[j + 3 * Qn], w3);
      const float2 s02 = make_float2(v0.x + v2.x, v0.y + v2.y), d02 = make_float2(v0.x - v2.x, v0.y - v2.y);
      const float2 s13 = make_float2(v1.x + v3.x, v1.y + v3.y), d13 = make_float2(v1.x - v3.x, v1.y - v3.y);
      const float2 jd = INVERSE ? make_float2(-d13.y, d13.x) : make_float2(d13.y, -d13.x);
      out[j0] = make_float2(s02.x + s13.x, s02.y + s13.y);
      out[j0 + Ns] = make_float2(d02.x + jd.x, d02.y + jd.y);
      out[j0 + 2 * Ns] = make_float2(s02.x - s13.x, s02.y - s13.y);
      out[j0 + 3 * Ns] = make_float2(d02.x - jd.x, d02.y - jd.y);
    }
    __syncwarp();
    float2 *t = in; in = out; out = t;
  }
  return in;
}

// Real transforms of length d through ONE complex transform of length h = d/2
// (z_m = x_{2m} + i x_{2m+1}):  with E, O the spectra of the even / odd samples,
//   Z_f = E_f + i O_f,  conj(Z_{h-f}) = E_f - i O_f,  X_f = E_f + W_d^f O_f,  X_h = E_0 - O_0.
// Per-warp scratch: two buffers of h float2 (= 2 d floats in total).
__host__ __device__ __forceinline__ size_t warp_fft_scratch_floats(int d) { return (size_t)2 * d; }

// packed spectrum (floats pk[0..d), in shared memory; may alias b1) -> time-domain row.
// The returned buffer, viewed as d floats, holds x_n * (d/2): scale by 2/d.
__device__ __forceinline__ const float *warp_irfft_packed(const float *pk, float2 *b0, float2 *b1, const float2 *tw,
                                                         int logd, int lane) {
  const int h = 1 << (logd - 1);
  __syncwarp();
  for (int f = lane; f < h; f += 32) {
    float2 z;
    if (f == 0) {
      const float x0 = pk[0], xh = pk[1];
      z = make_float2(0.5f * (x0 + xh), 0.5f * (x0 - xh));
    } else {
      const float2 xf = make_float2(pk[2 * f], pk[2 * f + 1]);
      const float2 xg = make_float2(pk[2 * (h - f)], -pk[2 * (h - f) + 1]);   // conj X_{h-f}
      const float2 e = make_float2(0.5f * (xf.x + xg.x), 0.5f * (xf.y + xg.y));
      const float2 dd = make_float2(0.5f * (xf.x - xg.x), 0.5f * (xf.y - xg.y));
      const float2 o = cmulc(tw[f], dd);                                       // W_d^{-f} * dd
      z = make_float2(e.x - o.y, e.y + o.x);                                   // E + i O
    }
    b0[f] = z;
  }
  __syncwarp();
  return reinterpret_cast<const float *>(warp_fft<true>(b0, b1, tw, 2, logd - 1, lane));
}

// b0 viewed as d floats holds the real row; returns Z (h float2), the half-length transform.
__device__ __forceinline__ const float2 *warp_rfft_half(float2 *b0, float2 *b1, const float2 *tw, int logd, int lane) {
  __syncwarp();
  return warp_fft<false>(b0, b1, tw, 2, logd - 1, lane);
}

// packed slot f (0 <= f < h) of the length-d spectrum from the half-length transform Z
__device__ __forceinline__ float2 packed_slot(const float2 *Z, int f, int h, const float2 *tw) {
  if (f == 0) return make_float2(Z[0].x + Z[0].y, Z[0].x - Z[0].y);   // (X_0, X_h)
  const float2 zf = Z[f], zg = Z[h - f];
  const float2 e = make_float2(0.5f * (zf.x + zg.x), 0.5f * (zf.y - zg.y));   // (Z_f + conj Z_g) / 2
  const float2 o = make_float2(0.5f * (zf.y + zg.y), -0.5f * (zf.x - zg.x));  // (Z_f - conj Z_g) / (2 i)
  const float2 wo = cmul(tw[f], o);
  return make_float2(e.x + wo.x, e.y + wo.y);
}

static inline int log2_exact(int d) {
  int l = 0;
  while ((1 << l) < d) ++l;
  return (1 << l) == d ? l : -1;
}

}  // namespace skge
