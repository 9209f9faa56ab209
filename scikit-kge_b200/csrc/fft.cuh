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

// One complex transform of length N = 1 << logd by ONE WARP: Stockham autosort, radix-4
// stages (plus a leading radix-2 stage when logd is odd), ping-ponging between `in` and
// `out` (both N float2 in shared memory, private to the warp).  Returns the buffer that
// holds the result.  Callers must __syncwarp() after filling `in`.
template <bool INVERSE>
__device__ __forceinline__ float2 *warp_fft(float2 *in, float2 *out, const float2 *tw, int logd, int lane) {
  const int N = 1 << logd, H = N >> 1, Qn = N >> 2;
  int Ns = 1;
  if (logd & 1) {
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
    const int tstep = N / (4 * Ns);
    for (int j = lane; j < Qn; j += 32) {
      const int k = j & (Ns - 1);
      const int j0 = ((j - k) << 2) + k;
      float2 w1 = tw_at(tw, k * tstep, H), w2 = tw_at(tw, 2 * k * tstep, H), w3 = tw_at(tw, 3 * k * tstep, H);
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

// Per-warp scratch: two complex buffers of N float2.
__host__ __device__ __forceinline__ size_t warp_fft_scratch_bytes(int d) { return (size_t)2 * d * sizeof(float2); }

// packed spectrum pk[0..d) (floats, shared or global, read through `ld`) -> time-domain row
// x[n] (unscaled: multiply by 1/d), left as the REAL parts of the returned buffer.
template <typename Load>
__device__ __forceinline__ float2 *warp_irfft_packed(Load ld, float2 *b0, float2 *b1, const float2 *tw, int logd,
                                                     int lane) {
  const int N = 1 << logd, H = N >> 1;
  __syncwarp();
  for (int f = lane; f < N; f += 32) {
    float2 v;
    if (f == 0) v = make_float2(ld(0), 0.f);
    else if (f == H) v = make_float2(ld(1), 0.f);
    else if (f < H) v = make_float2(ld(2 * f), ld(2 * f + 1));
    else v = make_float2(ld(2 * (N - f)), -ld(2 * (N - f) + 1));   // Hermitian extension
    b0[f] = v;
  }
  __syncwarp();
  return warp_fft<true>(b0, b1, tw, logd, lane);
}

// real row x[0..d) (read through `ld`) -> full complex spectrum in the returned buffer
// (only slots 0..d/2 are needed for the packed layout).
template <typename Load>
__device__ __forceinline__ float2 *warp_rfft(Load ld, float2 *b0, float2 *b1, const float2 *tw, int logd, int lane) {
  const int N = 1 << logd;
  __syncwarp();
  for (int n = lane; n < N; n += 32) b0[n] = make_float2(ld(n), 0.f);
  __syncwarp();
  return warp_fft<false>(b0, b1, tw, logd, lane);
}

// element p of the packed layout from a full spectrum X
__device__ __forceinline__ float packed_from_full(const float2 *X, int p, int H) {
  if (p == 0) return X[0].x;
  if (p == 1) return X[H].x;
  const float2 v = X[p >> 1];
  return (p & 1) ? v.y : v.x;
}

static inline int log2_exact(int d) {
  int l = 0;
  while ((1 << l) < d) ++l;
  return (1 << l) == d ? l : -1;
}

}  // namespace skge
