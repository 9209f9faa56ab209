// Model._scores(ss, ps, os) for TransE / HolE / RESCAL.
//   skge/transe.py:25-46, skge/hole.py:19-20, skge/rescal.py:31-35
#include "common.cuh"
#include "hole_math.cuh"

namespace skge {

// One warp per triple; rows gathered with VEC-wide loads.
template <int VEC>
__global__ void __launch_bounds__(256) transe_scores_kernel(const float *__restrict__ E,
                                                            const float *__restrict__ R,
                                                            const int32_t *__restrict__ s,
                                                            const int32_t *__restrict__ p,
                                                            const int32_t *__restrict__ o, int64_t n,
                                                            int d, int l1, float *__restrict__ out) {
  int lane = threadIdx.x & 31;
  int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t i = warp; i < n; i += nwarps) {
    const float *es = E + (int64_t)s[i] * d, *eo = E + (int64_t)o[i] * d, *rp = R + (int64_t)p[i] * d;
    float acc = 0.f;
    for (int c = lane * VEC; c < d; c += 32 * VEC) {
      float a[VEC], b[VEC], r[VEC];
      ld_vec<VEC>(es + c, a);
      ld_vec<VEC>(rp + c, r);
      ld_vec<VEC>(eo + c, b);
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        float x = a[v] + r[v] - b[v];
        acc += l1 ? fabsf(x) : x * x;
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) out[i] = -acc;
  }
}

// One CTA per triple: thread k owns ccorr(E[s],E[o])_k.
__global__ void hole_scores_kernel(const float *__restrict__ E, const float *__restrict__ R,
                                   const int32_t *__restrict__ s, const int32_t *__restrict__ p,
                                   const int32_t *__restrict__ o, int64_t n, int d,
                                   float *__restrict__ out) {
  extern __shared__ float sm[];
  float *a = sm, *o2 = sm + d, *red = sm + 3 * d;
  for (int64_t i = blockIdx.x; i < n; i += gridDim.x) {
    __syncthreads();
    smem_load(a, E + (int64_t)s[i] * d, d);
    smem_load_doubled(o2, E + (int64_t)o[i] * d, d);
    __syncthreads();
    const float *rp = R + (int64_t)p[i] * d;
    float part = 0.f;
    for (int k = threadIdx.x; k < d; k += blockDim.x) part += __ldg(rp + k) * sliding_dot(a, o2, k, d);
    float tot = block_sum(part, red);
    if (threadIdx.x == 0) out[i] = tot;
  }
}

// One CTA per triple: thread j owns (E[s]^T W[p])_j.
__global__ void rescal_scores_kernel(const float *__restrict__ E, const float *__restrict__ W,
                                     const int32_t *__restrict__ s, const int32_t *__restrict__ p,
                                     const int32_t *__restrict__ o, int64_t n, int d,
                                     float *__restrict__ out) {
  extern __shared__ float sm[];
  float *es = sm, *red = sm + d;
  for (int64_t i = blockIdx.x; i < n; i += gridDim.x) {
    __syncthreads();
    smem_load(es, E + (int64_t)s[i] * d, d);
    __syncthreads();
    const float *w = W + (int64_t)p[i] * d * d, *eo = E + (int64_t)o[i] * d;
    float part = 0.f;
    for (int j = threadIdx.x; j < d; j += blockDim.x) {
      float ew = 0.f;
      for (int r = 0; r < d; ++r) ew = fmaf(es[r], __ldg(w + (int64_t)r * d + j), ew);
      part += ew * __ldg(eo + j);
    }
    float tot = block_sum(part, red);
    if (threadIdx.x == 0) out[i] = tot;
  }
}

static int block_for(int d) {
  int t = (d + 31) / 32 * 32;
  return t < 64 ? 64 : (t > 256 ? 256 : t);
}

}  // namespace skge

using namespace skge;

extern "C" {

int skge_scores_transe(const float *E, const float *R, const int32_t *s, const int32_t *p,
                       const int32_t *o, int64_t n, int d, int l1, float *out, skge_stream_t stream) {
  SKGE_REQUIRE(d > 0 && n >= 0, "bad sizes");
  if (n == 0) return 0;
  int64_t blocks = (n + 7) / 8;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  cudaStream_t st = as_stream(stream);
  switch (pick_vec(d)) {
    case 4: transe_scores_kernel<4><<<(int)blocks, 256, 0, st>>>(E, R, s, p, o, n, d, l1, out); break;
    case 2: transe_scores_kernel<2><<<(int)blocks, 256, 0, st>>>(E, R, s, p, o, n, d, l1, out); break;
    default: transe_scores_kernel<1><<<(int)blocks, 256, 0, st>>>(E, R, s, p, o, n, d, l1, out); break;
  }
  SKGE_LAUNCH_CHECK();
  return 0;
}

int skge_scores_hole(const float *E, const float *R, const int32_t *s, const int32_t *p,
                     const int32_t *o, int64_t n, int d, float *out, skge_stream_t stream) {
  SKGE_REQUIRE(d > 0 && n >= 0 && d <= 8192, "bad sizes");
  if (n == 0) return 0;
  int64_t blocks = n > kNumSMs * 32 ? kNumSMs * 32 : n;
  size_t smem = (3 * (size_t)d + 40) * sizeof(float);
  SKGE_CUDA(cudaFuncSetAttribute(hole_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  hole_scores_kernel<<<(int)blocks, block_for(d), smem, as_stream(stream)>>>(E, R, s, p, o, n, d, out);
  SKGE_LAUNCH_CHECK();
  return 0;
}

int skge_scores_rescal(const float *E, const float *W, const int32_t *s, const int32_t *p,
                       const int32_t *o, int64_t n, int d, float *out, skge_stream_t stream) {
  SKGE_REQUIRE(d > 0 && n >= 0 && d <= 8192, "bad sizes");
  if (n == 0) return 0;
  int64_t blocks = n > kNumSMs * 32 ? kNumSMs * 32 : n;
  size_t smem = ((size_t)d + 40) * sizeof(float);
  rescal_scores_kernel<<<(int)blocks, block_for(d), smem, as_stream(stream)>>>(E, W, s, p, o, n, d, out);
  SKGE_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
