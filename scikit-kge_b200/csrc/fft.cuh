// Shared-memory FFT building blocks for HolE (even d whose half is 2^a 3^b 5^c).
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

// tw[m] = exp(-2 pi i m / N), m < N/2; index m < N by symmetry (N even, N/2 may be odd)
__device__ __forceinline__ float2 tw_at(const float2 *tw, int m, int half) {
  const bool up = m >= half;
  float2 w = tw[up ? m - half : m];
  return up ? make_float2(-w.x, -w.y) : w;
}

__device__ __forceinline__ void fill_twiddles(float2 *tw, int N, int tid, int nthreads) {
  for (int m = tid; m < N / 2; m += nthreads) {
    float sn, cs;
    sincospif(-2.0f * (float)m / (float)N, &sn, &cs);
    tw[m] = make_float2(cs, sn);
  }
}

// Lengths the warp transform handles: n = 2^a 3^b 5^c (numpy's FFT takes any length,
// skge/util.py:27,50; the configurations of the reference use d = 150 = 2 * 3 * 5 * 5).
__host__ __device__ __forceinline__ bool fft_len_ok(int n) {
  if (n < 1) return false;
  while ((n & 1) == 0) n >>= 1;
  while (n % 3 == 0) n /= 3;
  while (n % 5 == 0) n /= 5;
  return n == 1;
}
// packed spectra need an even row length whose half is such a length
__host__ __device__ __forceinline__ bool spectral_len_ok(int d) {
  return d >= 32 && d <= 1024 && (d & 1) == 0 && fft_len_ok(d / 2);
}

// One complex transform of length n = 2^a 3^b 5^c by ONE WARP: Stockham autosort, mixed radix:
// a leading radix-2 stage when a is odd, radix-4 stages for the rest of 2^a, then radix-3 and
// radix-5 stages, ping-ponging between `in` and `out` (n float2 each, shared memory private to
// the warp).  A radix-r stage with Ns = the product of the radices before it reads
// v_t = in[j + t n/r] W_{r Ns}^{t (j mod Ns)} and writes out[(j - j mod Ns) r + j mod Ns + s Ns] =
// sum_t v_t W_r^{s t}.  The twiddle table belongs to a transform of length n * tw_stride
// (tw[m] = exp(-2 pi i m / (n * tw_stride)), m < n * tw_stride / 2).  Returns the buffer that
// holds the result.  Callers must __syncwarp() after filling `in`.
template <bool INVERSE>
__device__ __forceinline__ float2 *warp_fft(float2 *in, float2 *out, const float2 *tw, int tw_stride, int n,
                                            int lane) {
  const int thalf = (n * tw_stride) >> 1;  // table entries
  int odd = n, log2n = 0;
  while ((odd & 1) == 0) { odd >>= 1; ++log2n; }
  const int N2 = n / odd;                  // the power-of-two part
  int Ns = 1;
  if (log2n & 1) {
    const int H = n >> 1;
    for (int j = lane; j < H; j += 32) {
      const float2 u0 = in[j], u1 = in[j + H];
      out[2 * j] = make_float2(u0.x + u1.x, u0.y + u1.y);
      out[2 * j + 1] = make_float2(u0.x - u1.x, u0.y - u1.y);
    }
    __syncwarp();
    float2 *t = in; in = out; out = t;
    Ns = 2;
  }
  const int Qn = n >> 2;
  for (; Ns < N2; Ns <<= 2) {
    const int tstep = (n / (4 * Ns)) * tw_stride;
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
  if (odd == 1) return in;   // power-of-two lengths end here
  for (int rest = odd; rest % 3 == 0; rest /= 3, Ns *= 3) {
    const int Tn = n / 3;
    const int tstep = (n / (3 * Ns)) * tw_stride;
    for (int j = lane; j < Tn; j += 32) {
      const int k = j % Ns;
      const int j0 = (j - k) * 3 + k;
      float2 w1 = tw_at(tw, k * tstep, thalf), w2 = tw_at(tw, 2 * k * tstep, thalf);
      if (INVERSE) { w1.y = -w1.y; w2.y = -w2.y; }
      const float2 v0 = in[j];
      const float2 v1 = cmul(in[j + Tn], w1);
      const float2 v2 = cmul(in[j + 2 * Tn], w2);
      const float2 s = make_float2(v1.x + v2.x, v1.y + v2.y);
      const float c = INVERSE ? -0.86602540378443865f : 0.86602540378443865f;
      const float2 e = make_float2(c * (v1.y - v2.y), -c * (v1.x - v2.x));   // -+ i sin(2 pi / 3) (v1 - v2)
      const float2 m = make_float2(fmaf(-0.5f, s.x, v0.x), fmaf(-0.5f, s.y, v0.y));
      out[j0] = make_float2(v0.x + s.x, v0.y + s.y);
      out[j0 + Ns] = make_float2(m.x + e.x, m.y + e.y);
      out[j0 + 2 * Ns] = make_float2(m.x - e.x, m.y - e.y);
    }
    __syncwarp();
    float2 *t = in; in = out; out = t;
  }
  for (int rest = odd; rest % 5 == 0; rest /= 5, Ns *= 5) {
    const int Fn = n / 5;
    const int tstep = (n / (5 * Ns)) * tw_stride;
    constexpr float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f;   // cos(2 pi / 5), cos(4 pi / 5)
    const float s1 = INVERSE ? -0.95105651629515357f : 0.95105651629515357f;  // sin(2 pi / 5)
    const float s2 = INVERSE ? -0.58778525229247313f : 0.58778525229247313f;  // sin(4 pi / 5)
    for (int j = lane; j < Fn; j += 32) {
      const int k = j % Ns;
      const int j0 = (j - k) * 5 + k;
      float2 w1 = tw_at(tw, k * tstep, thalf), w2 = tw_at(tw, 2 * k * tstep, thalf);
      float2 w3 = tw_at(tw, 3 * k * tstep, thalf), w4 = tw_at(tw, 4 * k * tstep, thalf);
      if (INVERSE) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; w4.y = -w4.y; }
      const float2 v0 = in[j];
      const float2 v1 = cmul(in[j + Fn], w1);
      const float2 v2 = cmul(in[j + 2 * Fn], w2);
      const float2 v3 = cmul(in[j + 3 * Fn], w3);
      const float2 v4 = cmul(in[j + 4 * Fn], w4);
      const float2 a1 = make_float2(v1.x + v4.x, v1.y + v4.y), b1 = make_float2(v1.x - v4.x, v1.y - v4.y);
      const float2 a2 = make_float2(v2.x + v3.x, v2.y + v3.y), b2 = make_float2(v2.x - v3.x, v2.y - v3.y);
      const float2 m1 = make_float2(v0.x + c1 * a1.x + c2 * a2.x, v0.y + c1 * a1.y + c2 * a2.y);
      const float2 m2 = make_float2(v0.x + c2 * a1.x + c1 * a2.x, v0.y + c2 * a1.y + c1 * a2.y);
      const float2 t1 = make_float2(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y);
      const float2 t2 = make_float2(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y);
      const float2 e1 = make_float2(t1.y, -t1.x), e2 = make_float2(t2.y, -t2.x);   // -+ i t (the sign is in s1, s2)
      out[j0] = make_float2(v0.x + a1.x + a2.x, v0.y + a1.y + a2.y);
      out[j0 + Ns] = make_float2(m1.x + e1.x, m1.y + e1.y);
      out[j0 + 2 * Ns] = make_float2(m2.x + e2.x, m2.y + e2.y);
      out[j0 + 3 * Ns] = make_float2(m2.x - e2.x, m2.y - e2.y);
      out[j0 + 4 * Ns] = make_float2(m1.x - e1.x, m1.y - e1.y);
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
                                                         int d, int lane) {
  const int h = d >> 1;
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
  return reinterpret_cast<const float *>(warp_fft<true>(b0, b1, tw, 2, h, lane));
}

// b0 viewed as d floats holds the real row; returns Z (h float2), the half-length transform.
__device__ __forceinline__ const float2 *warp_rfft_half(float2 *b0, float2 *b1, const float2 *tw, int d, int lane) {
  __syncwarp();
  return warp_fft<false>(b0, b1, tw, 2, d >> 1, lane);
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

// ---------------------------------------------------------------------------
// Register-resident transforms for d = 256 (one warp, no shared memory inside the transform).
// The half-length complex transform (128 points) is held four points per lane:
//   time side       z[j] of lane l  = z_m,  m = 32 j + l   (floats 64 j + 2 l, 64 j + 2 l + 1 of the row)
//   frequency side  Z[k] of lane p  = Z_f,  f = k + 4 brev5(p)
// 128 = 4 x 32: a radix-4 butterfly over the four registers, the twiddle W_128^{l k}, then a
// 32-point decimation-in-frequency transform ACROSS THE LANES (five __shfl_xor stages, which is
// what leaves the lane index bit-reversed).  The inverse runs the same graph backwards.
// The packed real spectrum (see the top of this file) is formed in the same permuted order;
// the partner slot 128 - f lives in lane 31 - p, register 4 - k (k > 0) or lane q0, register 0.
// ---------------------------------------------------------------------------
struct RegFft256 {
  float2 w1, w2, w3;   // W_128^{l k}, k = 1..3
  float2 ws[4];        // lane stages of span 16, 8, 4, 2: W_{2 span}^{l mod span} in the upper lane, 1 in the lower
  float sg[5];         // -1 in the upper lane of a stage, +1 in the lower (spans 16, 8, 4, 2, 1)
  float2 wrh[4];       // W_256^f / 2 for the lane's four slots
  int q0;              // lane holding slot 128 - 4 brev5(p) in register 0
};

__device__ __forceinline__ int brev5(int x) { return (int)(__brev((unsigned)x) >> 27); }

__device__ __forceinline__ float2 unit_root(float num, float den) {   // exp(-2 pi i num / den)
  float sn, cs;
  sincospif(-2.0f * num / den, &sn, &cs);
  return make_float2(cs, sn);
}

__device__ __forceinline__ void regfft256_init(RegFft256 &c, int lane) {
  c.w1 = unit_root((float)lane, 128.f);
  c.w2 = unit_root((float)(2 * lane), 128.f);
  c.w3 = unit_root((float)(3 * lane), 128.f);
#pragma unroll
  for (int si = 0; si < 5; ++si) {
    const int s = 16 >> si;
    const bool upper = (lane & s) != 0;
    c.sg[si] = upper ? -1.f : 1.f;
    if (si < 4) c.ws[si] = upper ? unit_root((float)(lane & (s - 1)), (float)(2 * s)) : make_float2(1.f, 0.f);
  }
  const int bp = brev5(lane);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 w = unit_root((float)(k + 4 * bp), 256.f);
    c.wrh[k] = make_float2(0.5f * w.x, 0.5f * w.y);
  }
  c.q0 = brev5((32 - bp) & 31);
}

// z (time order) -> Z (frequency order), unnormalised
__device__ __forceinline__ void regfft256_fwd(float2 (&z)[4], const RegFft256 &c) {
  {
    const float2 s02 = make_float2(z[0].x + z[2].x, z[0].y + z[2].y), d02 = make_float2(z[0].x - z[2].x, z[0].y - z[2].y);
    const float2 s13 = make_float2(z[1].x + z[3].x, z[1].y + z[3].y), d13 = make_float2(z[1].x - z[3].x, z[1].y - z[3].y);
    z[0] = make_float2(s02.x + s13.x, s02.y + s13.y);
    z[1] = cmul(make_float2(d02.x + d13.y, d02.y - d13.x), c.w1);   // d02 - i d13
    z[2] = cmul(make_float2(s02.x - s13.x, s02.y - s13.y), c.w2);
    z[3] = cmul(make_float2(d02.x - d13.y, d02.y + d13.x), c.w3);   // d02 + i d13
  }
#pragma unroll
  for (int si = 0; si < 5; ++si) {
    const int s = 16 >> si;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float ox = __shfl_xor_sync(kFull, z[k].x, s), oy = __shfl_xor_sync(kFull, z[k].y, s);
      const float2 u = make_float2(fmaf(c.sg[si], z[k].x, ox), fmaf(c.sg[si], z[k].y, oy));
      z[k] = si < 4 ? cmul(u, c.ws[si]) : u;
    }
  }
}

// Z (frequency order) -> z (time order) times 128
__device__ __forceinline__ void regfft256_inv(float2 (&z)[4], const RegFft256 &c) {
#pragma unroll
  for (int si = 4; si >= 0; --si) {
    const int s = 16 >> si;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 t = si < 4 ? cmulc(c.ws[si], z[k]) : z[k];
      const float ox = __shfl_xor_sync(kFull, t.x, s), oy = __shfl_xor_sync(kFull, t.y, s);
      z[k] = make_float2(fmaf(c.sg[si], t.x, ox), fmaf(c.sg[si], t.y, oy));
    }
  }
  const float2 y1 = cmulc(c.w1, z[1]), y2 = cmulc(c.w2, z[2]), y3 = cmulc(c.w3, z[3]);
  const float2 s02 = make_float2(z[0].x + y2.x, z[0].y + y2.y), d02 = make_float2(z[0].x - y2.x, z[0].y - y2.y);
  const float2 s13 = make_float2(y1.x + y3.x, y1.y + y3.y), d13 = make_float2(y1.x - y3.x, y1.y - y3.y);
  z[0] = make_float2(s02.x + s13.x, s02.y + s13.y);
  z[1] = make_float2(d02.x - d13.y, d02.y + d13.x);   // d02 + i d13
  z[2] = make_float2(s02.x - s13.x, s02.y - s13.y);
  z[3] = make_float2(d02.x + d13.y, d02.y - d13.x);   // d02 - i d13
}

// the four partner values (slot 128 - f) of a lane's registers
__device__ __forceinline__ void regfft256_partners(const float2 (&v)[4], float2 (&g)[4], const RegFft256 &c, int lane) {
  const int m = 31 - lane;
  g[0] = make_float2(__shfl_sync(kFull, v[0].x, c.q0), __shfl_sync(kFull, v[0].y, c.q0));
  g[1] = make_float2(__shfl_sync(kFull, v[3].x, m), __shfl_sync(kFull, v[3].y, m));
  g[2] = make_float2(__shfl_sync(kFull, v[2].x, m), __shfl_sync(kFull, v[2].y, m));
  g[3] = make_float2(__shfl_sync(kFull, v[1].x, m), __shfl_sync(kFull, v[1].y, m));
}

// packed real spectrum X (frequency order) -> real row: z[j] = (x_{64 j + 2 l}, x_{64 j + 2 l + 1}) * 128
__device__ __forceinline__ void regfft256_irfft(float2 (&v)[4], const RegFft256 &c, int lane) {
  float2 g[4];
  regfft256_partners(v, g, c, lane);
  const float2 z00 = make_float2(0.5f * (v[0].x + v[0].y), 0.5f * (v[0].x - v[0].y));   // slot 0 = (X_0, X_128)
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 e2 = make_float2(v[k].x + g[k].x, v[k].y - g[k].y);     // X_f + conj X_{128-f}
    const float2 d2 = make_float2(v[k].x - g[k].x, v[k].y + g[k].y);     // X_f - conj X_{128-f}
    const float2 o = cmulc(c.wrh[k], d2);                                // W_256^{-f} (X_f - conj X_{128-f}) / 2
    v[k] = make_float2(fmaf(0.5f, e2.x, -o.y), fmaf(0.5f, e2.y, o.x));   // E_f + i O_f
  }
  if (lane == 0) v[0] = z00;
  regfft256_inv(v, c);
}

// real row (time order, as above) -> packed real spectrum (frequency order)
__device__ __forceinline__ void regfft256_rfft(float2 (&v)[4], const RegFft256 &c, int lane) {
  regfft256_fwd(v, c);
  float2 g[4];
  regfft256_partners(v, g, c, lane);
  const float2 x00 = make_float2(v[0].x + v[0].y, v[0].x - v[0].y);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 e2 = make_float2(v[k].x + g[k].x, v[k].y - g[k].y);     // Z_f + conj Z_{128-f}
    const float2 o2 = make_float2(v[k].y + g[k].y, g[k].x - v[k].x);     // -i (Z_f - conj Z_{128-f})
    const float2 t = cmul(c.wrh[k], o2);
    v[k] = make_float2(fmaf(0.5f, e2.x, t.x), fmaf(0.5f, e2.y, t.y));
  }
  if (lane == 0) v[0] = x00;
}

// Transposition between the frequency order above and the packed row in memory order (lane l owns
// the 16-byte units l and 32 + l) through 1 KB of per-warp shared memory; units are XOR-swizzled so
// that both access patterns are conflict-free.
__device__ __forceinline__ int regfft256_swz(int u) { return u ^ ((u >> 3) & 7); }

__device__ __forceinline__ void regfft256_row_to_freq(float4 *buf, const float4 &r0, const float4 &r1, float2 (&v)[4], int lane) {
  __syncwarp();
  buf[regfft256_swz(lane)] = r0;
  buf[regfft256_swz(32 + lane)] = r1;
  __syncwarp();
  const int u = 2 * brev5(lane);
  const float4 t0 = buf[regfft256_swz(u)], t1 = buf[regfft256_swz(u + 1)];
  v[0] = make_float2(t0.x, t0.y); v[1] = make_float2(t0.z, t0.w);
  v[2] = make_float2(t1.x, t1.y); v[3] = make_float2(t1.z, t1.w);
}

__device__ __forceinline__ void regfft256_freq_to_row(float4 *buf, const float2 (&v)[4], float4 &r0, float4 &r1, int lane) {
  __syncwarp();
  const int u = 2 * brev5(lane);
  buf[regfft256_swz(u)] = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
  buf[regfft256_swz(u + 1)] = make_float4(v[2].x, v[2].y, v[3].x, v[3].y);
  __syncwarp();
  r0 = buf[regfft256_swz(lane)];
  r1 = buf[regfft256_swz(32 + lane)];
}

static inline int log2_exact(int d) {
  int l = 0;
  while ((1 << l) < d) ++l;
  return (1 << l) == d ? l : -1;
}

}  // namespace skge
