// Shared device/host helpers for libskge_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "skge_b200.h"

namespace skge {

// ---- error plumbing -------------------------------------------------------
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define SKGE_CUDA(call)                                                        \
  do {                                                                         \
    cudaError_t _e = (call);                                                   \
    if (_e != cudaSuccess) return ::skge::cuda_fail(_e, #call, __FILE__, __LINE__); \
  } while (0)

#define SKGE_LAUNCH_CHECK() SKGE_CUDA(cudaPeekAtLastError())

#define SKGE_REQUIRE(cond, msg)                      \
  do {                                               \
    if (!(cond)) {                                   \
      ::skge::set_error("%s: %s", __func__, msg);    \
      return SKGE_EINVAL;                            \
    }                                                \
  } while (0)

static inline cudaStream_t as_stream(skge_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Bump allocator over the caller's workspace.
struct Arena {
  char *base;
  size_t cap, off;
  Arena(void *p, size_t bytes) : base(static_cast<char *>(p)), cap(bytes), off(0) {}
  template <typename T>
  T *take(size_t n) {
    size_t bytes = align_up(n * sizeof(T));
    T *r = reinterpret_cast<T *>(base + off);
    off += bytes;
    return r;
  }
  bool ok() const { return off <= cap; }
};

static constexpr int kNumSMs = 148;  // B200

// ---- device helpers ---------------------------------------------------------
#ifdef __CUDACC__
static constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// Rows are d floats; a row start is 16-byte aligned iff (row * d) % 4 == 0, which
// holds for every row when d % 4 == 0.  VEC is chosen on the host from d.
template <int VEC>
struct VecT;
template <>
struct VecT<4> { using type = float4; };
template <>
struct VecT<2> { using type = float2; };
template <>
struct VecT<1> { using type = float; };

template <int VEC>
__device__ __forceinline__ void ld_vec(const float *p, float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    float4 t = __ldg(reinterpret_cast<const float4 *>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else if constexpr (VEC == 2) {
    float2 t = __ldg(reinterpret_cast<const float2 *>(p));
    v[0] = t.x; v[1] = t.y;
  } else {
    v[0] = __ldg(p);
  }
}
// plain (coherent) load: for tables that the same kernel also writes
template <int VEC>
__device__ __forceinline__ void ld_vec_rw(const float *p, float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    float4 t = *reinterpret_cast<const float4 *>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else if constexpr (VEC == 2) {
    float2 t = *reinterpret_cast<const float2 *>(p);
    v[0] = t.x; v[1] = t.y;
  } else {
    v[0] = *p;
  }
}
template <int VEC>
__device__ __forceinline__ void st_vec(float *p, const float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
  } else if constexpr (VEC == 2) {
    *reinterpret_cast<float2 *>(p) = make_float2(v[0], v[1]);
  } else {
    *p = v[0];
  }
}

// AdaGrad's 1 / max(sqrt(p2), 1e-7) (skge/param.py:152-155) as one MUFU.RSQ (2 ulp) instead of
// an IEEE sqrt followed by an IEEE division: the row update is issue-bound, not byte-bound.
__device__ __forceinline__ float adagrad_rscale(float p2) { return fminf(rsqrtf(p2), 1e7f); }

__device__ __forceinline__ float act_f(int af, float x) {
  switch (af) {
    case SKGE_AF_SIGMOID: return 1.0f / (1.0f + expf(-x));
    case SKGE_AF_TANH: return tanhf(x);
    case SKGE_AF_RELU: return fmaxf(0.0f, x);
    default: return x;
  }
}
__device__ __forceinline__ float act_g_given_f(int af, float fx) {
  switch (af) {
    case SKGE_AF_SIGMOID: return fx * (1.0f - fx);
    case SKGE_AF_TANH: return 1.0f - fx * fx;
    case SKGE_AF_RELU: return fx > 0.0f ? 1.0f : 0.0f;
    default: return 1.0f;
  }
}
#endif  // __CUDACC__

static inline int pick_vec(int64_t rowlen) { return (rowlen % 4 == 0) ? 4 : (rowlen % 2 == 0) ? 2 : 1; }

}  // namespace skge
