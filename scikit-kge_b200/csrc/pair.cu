// Pairwise-margin minibatch kernels (TransE, HolE) and their C entry points.
//   TransE._pairwise_gradients : skge/transe.py:48-165
//   HolE._pairwise_gradients   : skge/hole.py:44-100
//   PairwiseStochasticTrainer._process_batch / _batch_step : skge/base.py:1394-1427, 1306-1316
#include "common.cuh"
#include "fft.cuh"
#include "hole_math.cuh"
#include "segment.cuh"
#include "umma.cuh"

namespace skge {

struct PairIdx {
  const int32_t *sp, *op, *pp, *sn, *on, *pn;
  const uint8_t *valid;
};

// ---------------------------------------------------------------------------
// TransE: one warp per pair.  Pass 1 gathers the six rows and reduces both
// distances; pass 2 (violating pairs only, rows now in L1) writes the two
// per-pair gradient rows pg, ng.
//   G[i][0] = pg = sign(E[sp]+R[pp]-E[op])  (L1)  |  E[sp]+R[pp]-E[op]      (L2)
//   G[i][1] = ng = sign(E[on]-R[pn]-E[sn])  (L1)  |  E[on]-R[pn]-E[sn]      (L2)
// (skge/transe.py:103-121; note np.sign(0) == 0 and no factor 2 for L2.)
// ---------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(256) transe_pair_kernel(const float *__restrict__ E,
                                                          const float *__restrict__ R, PairIdx ix, int64_t P,
                                                          int d, int l1, float margin,
                                                          float *__restrict__ pscores,
                                                          float *__restrict__ nscores,
                                                          uint8_t *__restrict__ flags, float *__restrict__ G,
                                                          int32_t *__restrict__ counts,
                                                          int64_t *__restrict__ nviol_accum,
                                                          int32_t *__restrict__ ent_viol) {
  const int lane = threadIdx.x & 31;
  int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  int nv = 0;
  for (int64_t i = warp; i < P; i += nwarps) {
    if (ix.valid && !ix.valid[i]) {
      if (lane == 0) {
        flags[i] = 0;
        if (pscores) pscores[i] = 0.f;
        if (nscores) nscores[i] = 0.f;
      }
      continue;
    }
    int sp = ix.sp[i], op = ix.op[i], pp = ix.pp[i], sn = ix.sn[i], on = ix.on[i], pn = ix.pn[i];
    const float *esp = E + (int64_t)sp * d, *eop = E + (int64_t)op * d, *rpp = R + (int64_t)pp * d;
    const float *esn = E + (int64_t)sn * d, *eon = E + (int64_t)on * d, *rpn = R + (int64_t)pn * d;
    float ap = 0.f, an = 0.f;
    for (int c = lane * VEC; c < d; c += 32 * VEC) {
      float a[VEC], b[VEC], r[VEC], a2[VEC], b2[VEC], r2[VEC];
      ld_vec<VEC>(esp + c, a);
      ld_vec<VEC>(rpp + c, r);
      ld_vec<VEC>(eop + c, b);
      ld_vec<VEC>(esn + c, a2);
      ld_vec<VEC>(rpn + c, r2);
      ld_vec<VEC>(eon + c, b2);
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        float x = a[v] + r[v] - b[v];
        float y = a2[v] + r2[v] - b2[v];
        ap += l1 ? fabsf(x) : x * x;
        an += l1 ? fabsf(y) : y * y;
      }
    }
    float ps = -warp_sum(ap), ns = -warp_sum(an);
    bool viol = ns + margin > ps;  // skge/transe.py:73
    if (lane == 0) {
      flags[i] = viol;
      if (pscores) pscores[i] = ps;
      if (nscores) nscores[i] = ns;
    }
    if (!viol) continue;
    ++nv;
    if (ent_viol && lane == 0) {  // distinct entities of the pair: skge/transe.py:78-83
      atomicAdd(ent_viol + sn, 1);
      if (on != sn) atomicAdd(ent_viol + on, 1);
      if (sp != sn && sp != on) atomicAdd(ent_viol + sp, 1);
      if (op != sn && op != on && op != sp) atomicAdd(ent_viol + op, 1);
    }
    float *gp = G + (int64_t)i * 2 * d, *gn = gp + d;
    for (int c = lane * VEC; c < d; c += 32 * VEC) {
      float a[VEC], b[VEC], r[VEC], a2[VEC], b2[VEC], r2[VEC], o1[VEC], o2[VEC];
      ld_vec<VEC>(esp + c, a);
      ld_vec<VEC>(rpp + c, r);
      ld_vec<VEC>(eop + c, b);
      ld_vec<VEC>(esn + c, a2);
      ld_vec<VEC>(rpn + c, r2);
      ld_vec<VEC>(eon + c, b2);
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        float x = a[v] + r[v] - b[v];     // -(E[op]-R[pp]-E[sp])
        float y = b2[v] - r2[v] - a2[v];  //   E[on]-R[pn]-E[sn]
        if (l1) {
          x = (float)((x > 0.f) - (x < 0.f));
          y = (float)((y > 0.f) - (y < 0.f));
        }
        o1[v] = x;
        o2[v] = y;
      }
      st_vec<VEC>(gp + c, o1);
      st_vec<VEC>(gn + c, o2);
    }
  }
  if (lane == 0 && nv) {
    atomicAdd(counts, nv);
    if (nviol_accum) atomicAdd(reinterpret_cast<unsigned long long *>(nviol_accum), (unsigned long long)nv);
  }
}

// Same, for rows of at most MAXC * 32 * VEC floats: the two difference vectors stay in registers
// between the score and the gradient rows, so every input row is requested once.
__device__ __forceinline__ float sign0(float x) { return copysignf(x != 0.f ? 1.f : 0.f, x); }  // np.sign

template <int VEC, int MAXC>
__global__ void __launch_bounds__(256, 4) transe_pair_reg_kernel(const float *__restrict__ E,
                                                                 const float *__restrict__ R, PairIdx ix,
                                                                 int64_t P, int d, int l1, float margin,
                                                                 float *__restrict__ pscores,
                                                                 float *__restrict__ nscores,
                                                                 uint8_t *__restrict__ flags,
                                                                 float *__restrict__ G,
                                                                 int32_t *__restrict__ counts,
                                                                 int64_t *__restrict__ nviol_accum,
                                                                 int32_t *__restrict__ ent_viol) {
  const int lane = threadIdx.x & 31;
  int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  int nv = 0;
  for (int64_t i = warp; i < P; i += nwarps) {
    if (ix.valid && !ix.valid[i]) {
      if (lane == 0) {
        flags[i] = 0;
        if (pscores) pscores[i] = 0.f;
        if (nscores) nscores[i] = 0.f;
      }
      continue;
    }
    int sp = ix.sp[i], op = ix.op[i], pp = ix.pp[i], sn = ix.sn[i], on = ix.on[i], pn = ix.pn[i];
    const float *esp = E + (int64_t)sp * d, *eop = E + (int64_t)op * d, *rpp = R + (int64_t)pp * d;
    const float *esn = E + (int64_t)sn * d, *eon = E + (int64_t)on * d, *rpn = R + (int64_t)pn * d;
    float x[MAXC][VEC], y[MAXC][VEC];  // E[s]+R[p]-E[o] of the positive / negative triple
    float ap = 0.f, an = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int col = (c * 32 + lane) * VEC;
      if (col < d) {
        float a[VEC], b[VEC], r[VEC], a2[VEC], b2[VEC], r2[VEC];
        ld_vec<VEC>(esp + col, a);
        ld_vec<VEC>(rpp + col, r);
        ld_vec<VEC>(eop + col, b);
        ld_vec<VEC>(esn + col, a2);
        ld_vec<VEC>(rpn + col, r2);
        ld_vec<VEC>(eon + col, b2);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          x[c][v] = a[v] + r[v] - b[v];
          y[c][v] = a2[v] + r2[v] - b2[v];
          ap += l1 ? fabsf(x[c][v]) : x[c][v] * x[c][v];
          an += l1 ? fabsf(y[c][v]) : y[c][v] * y[c][v];
        }
      }
    }
    float ps = -warp_sum(ap), ns = -warp_sum(an);
    bool viol = ns + margin > ps;  // skge/transe.py:73
    if (lane == 0) {
      flags[i] = viol;
      if (pscores) pscores[i] = ps;
      if (nscores) nscores[i] = ns;
    }
    if (!viol) continue;
    ++nv;
    if (ent_viol && lane == 0) {  // distinct entities of the pair: skge/transe.py:78-83
      atomicAdd(ent_viol + sn, 1);
      if (on != sn) atomicAdd(ent_viol + on, 1);
      if (sp != sn && sp != on) atomicAdd(ent_viol + sp, 1);
      if (op != sn && op != on && op != sp) atomicAdd(ent_viol + op, 1);
    }
    float *gp = G + (int64_t)i * 2 * d, *gn = gp + d;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int col = (c * 32 + lane) * VEC;
      if (col < d) {
        float o1[VEC], o2[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          o1[v] = l1 ? sign0(x[c][v]) : x[c][v];     // -(E[op]-R[pp]-E[sp])
          o2[v] = l1 ? -sign0(y[c][v]) : -y[c][v];   //   E[on]-R[pn]-E[sn]
        }
        st_vec<VEC>(gp + col, o1);
        st_vec<VEC>(gn + col, o2);
      }
    }
  }
  if (lane == 0 && nv) {
    atomicAdd(counts, nv);
    if (nviol_accum) atomicAdd(reinterpret_cast<unsigned long long *>(nviol_accum), (unsigned long long)nv);
  }
}

// ---------------------------------------------------------------------------
// HolE: one CTA per pair, thread k owns component k of every correlation.
// Phase 1: ccorr(s,o) for the positive and the negative triple -> raw scores
// -> activation -> margin test.  Phase 2 (violators): the remaining four
// correlations, scaled by -/+ g_given_f and written as six rows:
//   G[i][0] = gp ccorr(R[pp],E[op]) -> sp     G[i][1] = gn ccorr(R[pn],E[on]) -> sn
//   G[i][2] = gp cconv(E[sp],R[pp]) -> op     G[i][3] = gn cconv(E[sn],R[pn]) -> on
//   G[i][4] = gp ccorr(E[sp],E[op]) -> pp     G[i][5] = gn ccorr(E[sn],E[on]) -> pn
// (skge/hole.py:66-97)
// ---------------------------------------------------------------------------
// Rows (2q, 2q + 1) of a pair go to the same slot (subject / object / relation) of the positive
// and the negative triple.  When both triples hold the same id there, the two contributions are
// summed here and only row 2q is used (RoleMap::twin): a corrupted pair shares two of its three
// slots, so it emits four gradient rows instead of six.
template <typename T>
__device__ __forceinline__ void store_folded(T *row, int stride, bool same, T vp, T vn);
template <>
__device__ __forceinline__ void store_folded<float>(float *row, int stride, bool same, float vp, float vn) {
  if (same) {
    row[0] = vp + vn;
  } else {
    row[0] = vp;
    row[stride] = vn;
  }
}
template <>
__device__ __forceinline__ void store_folded<float2>(float2 *row, int stride, bool same, float2 vp, float2 vn) {
  if (same) {
    row[0] = make_float2(vp.x + vn.x, vp.y + vn.y);
  } else {
    row[0] = vp;
    row[stride] = vn;
  }
}

// Register-blocked (hole_math.cuh): the CTA is four groups of d4/4 threads, thread j of a group
// owns outputs 4j..4j+3 of the group's correlation.
//   phase 1: groups (0, 2) / (1, 3) each take half of the input range of ccorr(s, o) of the
//            positive / negative triple; the score is linear in the partial results.
//   phase 2: groups 0..3 take ccorr(r,o)+, ccorr(r,o)-, cconv(s,r)+, cconv(s,r)-; the six scaled
//            rows are staged in shared memory and written out folded and coalesced.
__global__ void hole_pair_kernel(const float *__restrict__ E, const float *__restrict__ R, PairIdx ix,
                                 int64_t P, int d, int af, float margin, float *__restrict__ pscores,
                                 float *__restrict__ nscores, uint8_t *__restrict__ flags,
                                 float *__restrict__ G, int32_t *__restrict__ counts,
                                 int64_t *__restrict__ nviol_accum) {
  extern __shared__ __align__(16) float sm[];
  const int d4 = round4(d), T4 = d4 >> 2;
  // per triple: s[d4], srev[d4], r[d4], o2[2 d4], r2[2 d4]; then the staging rows [6][d4]
  float *tri[2];
  tri[0] = sm;
  tri[1] = sm + 7 * d4;
  float *stage = sm + 14 * d4;
  float *red = stage + 6 * d4;
  const int grp = threadIdx.x / T4, j = threadIdx.x - grp * T4, k0 = 4 * j;
  const bool worker = grp < 4;
  for (int64_t i = blockIdx.x; i < P; i += gridDim.x) {
    if (ix.valid && !ix.valid[i]) {
      if (threadIdx.x == 0) {
        flags[i] = 0;
        if (pscores) pscores[i] = 0.f;
        if (nscores) nscores[i] = 0.f;
      }
      continue;
    }
    __syncthreads();
    const int sid[2] = {ix.sp[i], ix.sn[i]}, oid[2] = {ix.op[i], ix.on[i]}, pid[2] = {ix.pp[i], ix.pn[i]};
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const float *es = E + (int64_t)sid[t] * d, *eo = E + (int64_t)oid[t] * d, *rr = R + (int64_t)pid[t] * d;
      smem_load_padded(tri[t], es, d, d4);
      smem_load_rev_padded(tri[t] + d4, es, d, d4);
      smem_load_padded(tri[t] + 2 * d4, rr, d, d4);
      smem_load_periodic(tri[t] + 3 * d4, eo, d, d4);
      smem_load_periodic(tri[t] + 5 * d4, rr, d, d4);
    }
    __syncthreads();
    // phase 1
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    float part = 0.f;
    const int t1 = grp & 1;
    const float *mine = sm + t1 * 7 * d4;  // this group's triple
    if (worker) {
      const int ih = ((T4 + 1) >> 1) << 2;  // first half of the input range, a multiple of 4
      const int ib = (grp >> 1) ? ih : 0, ie = (grp >> 1) ? d4 : ih;
      sliding_dot4(mine, mine + 3 * d4, k0, ib, ie, c);
      const float *rv = mine + 2 * d4;  // zero beyond d, so garbage outputs k >= d drop out
#pragma unroll
      for (int m = 0; m < 4; ++m) part += rv[k0 + m] * c[m];
    }
    const float raw_p = block_sum(worker && t1 == 0 ? part : 0.f, red);
    const float raw_n = block_sum(worker && t1 == 1 ? part : 0.f, red);
    const float fp = act_f(af, raw_p), fn = act_f(af, raw_n);
    const bool viol = fn + margin > fp;  // skge/hole.py:56
    if (threadIdx.x == 0) {
      flags[i] = viol;
      if (pscores) pscores[i] = raw_p;
      if (nscores) nscores[i] = raw_n;
      if (viol) {
        atomicAdd(counts, 1);
        if (nviol_accum) atomicAdd(reinterpret_cast<unsigned long long *>(nviol_accum), 1ull);
      }
    }
    if (!viol) continue;  // uniform over the CTA
    const float gs[2] = {-act_g_given_f(af, fp), act_g_given_f(af, fn)};  // skge/hole.py:66-67
    // rows 4, 5 (-> pp, pn): g * ccorr(s, o); the two halves meet in the staging row
    if (worker && (grp >> 1) == 1)
      *reinterpret_cast<float4 *>(stage + (4 + t1) * d4 + k0) = make_float4(c[0], c[1], c[2], c[3]);
    __syncthreads();
    if (worker && (grp >> 1) == 0) {
      float4 *dst = reinterpret_cast<float4 *>(stage + (4 + t1) * d4 + k0);
      const float4 o = *dst;
      const float g = gs[t1];
      *dst = make_float4(g * (c[0] + o.x), g * (c[1] + o.y), g * (c[2] + o.z), g * (c[3] + o.w));
    }
    // phase 2: rows 0, 1 (-> sp, sn): g * ccorr(r, o); rows 2, 3 (-> op, on): g * cconv(s, r)
    if (worker) {
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      const float *a = (grp >> 1) ? mine + d4 : mine + 2 * d4;        // rev(s) | r
      const float *x2 = (grp >> 1) ? mine + 5 * d4 : mine + 3 * d4;  // r2 | o2
      sliding_dot4(a, x2, k0, 0, d4, v);
      const float g = gs[t1];
      *reinterpret_cast<float4 *>(stage + grp * d4 + k0) = make_float4(g * v[0], g * v[1], g * v[2], g * v[3]);
    }
    __syncthreads();
    float *g = G + (int64_t)i * 6 * d;
    const bool same[3] = {sid[0] == sid[1], oid[0] == oid[1], pid[0] == pid[1]};
    for (int e = threadIdx.x; e < 3 * d; e += blockDim.x) {
      const int q = e / d, k = e - q * d;
      store_folded(g + 2 * q * d + k, d, same[q], stage[2 * q * d4 + k], stage[(2 * q + 1) * d4 + k]);
    }
  }
}

// ---------------------------------------------------------------------------
// HolE for power-of-two d: the correlations go through radix-2 Stockham FFTs in shared
// memory (what the reference does with numpy's FFT, skge/util.py:27,50), O(d log d)
// instead of O(d^2) per correlation.  Per pair, three complex FFTs carry the six real
// rows (s + i o, r + i r', s' + i o'); spectra are unpacked by Hermitian symmetry, the
// six products formed, the raw scores read off by Parseval
//     score = sum_k r_k ccorr(s,o)_k = (1/d) Re sum_f conj(S_f) O_f conj(R_f),
// and, for violating pairs only, three inverse FFTs return the six gradient rows
// (two real rows per complex transform).
// ---------------------------------------------------------------------------
// NF transforms of length N = 1 << LOGD by a Stockham autosort FFT: radix-4 stages (plus one
// leading radix-2 stage when LOGD is odd), ping-ponging between `in` and `out`; returns the
// buffer that holds the result.  tw[m] = exp(-2 pi i m / N), m < N/2.  One butterfly index
// per thread and stage is shared by the NF transforms (same twiddles, same addresses).
template <int LOGD, int NF, bool INVERSE>
__device__ __forceinline__ float2 *fft_batch(float2 *in, float2 *out, const float2 *tw) {
  constexpr int N = 1 << LOGD, H = N / 2, Qn = N / 4;
  int Ns = 1;
  if (LOGD & 1) {  // radix-2 stage with Ns = 1: no twiddles
    for (int j = threadIdx.x; j < H; j += blockDim.x) {
#pragma unroll
      for (int f = 0; f < NF; ++f) {
        const float2 u0 = in[f * N + j], u1 = in[f * N + j + H];
        out[f * N + 2 * j] = make_float2(u0.x + u1.x, u0.y + u1.y);
        out[f * N + 2 * j + 1] = make_float2(u0.x - u1.x, u0.y - u1.y);
      }
    }
    __syncthreads();
    float2 *t = in; in = out; out = t;
    Ns = 2;
  }
#pragma unroll 1
  for (; Ns < N; Ns <<= 2) {
    const int tstep = N / (4 * Ns);  // twiddle index of exp(-2 pi i k / (4 Ns)) is k * tstep
    for (int j = threadIdx.x; j < Qn; j += blockDim.x) {
      const int k = j & (Ns - 1);
      const int j0 = ((j - k) << 2) + k;
      float2 w1 = tw_at(tw, k * tstep, H), w2 = tw_at(tw, 2 * k * tstep, H), w3 = tw_at(tw, 3 * k * tstep, H);
      if (INVERSE) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; }
#pragma unroll
      for (int f = 0; f < NF; ++f) {
        const float2 v0 = in[f * N + j];
        const float2 v1 = cmul(in[f * N + j + Qn], w1);
        const float2 v2 = cmul(in[f * N + j + 2 * Qn], w2);
        const float2 v3 = cmul(in[f * N + j + 3 * Qn], w3);
        const float2 s02 = make_float2(v0.x + v2.x, v0.y + v2.y), d02 = make_float2(v0.x - v2.x, v0.y - v2.y);
        const float2 s13 = make_float2(v1.x + v3.x, v1.y + v3.y), d13 = make_float2(v1.x - v3.x, v1.y - v3.y);
        // forward: y1 = d02 - i d13, y3 = d02 + i d13 ; inverse: the other way round
        const float2 jd = INVERSE ? make_float2(-d13.y, d13.x) : make_float2(d13.y, -d13.x);
        out[f * N + j0] = make_float2(s02.x + s13.x, s02.y + s13.y);
        out[f * N + j0 + Ns] = make_float2(d02.x + jd.x, d02.y + jd.y);
        out[f * N + j0 + 2 * Ns] = make_float2(s02.x - s13.x, s02.y - s13.y);
        out[f * N + j0 + 3 * Ns] = make_float2(d02.x - jd.x, d02.y - jd.y);
      }
    }
    __syncthreads();
    float2 *t = in; in = out; out = t;
  }
  return in;
}

template <int LOGD>
__global__ void __launch_bounds__(256) hole_pair_fft_kernel(const float *__restrict__ E, const float *__restrict__ R,
                                                            PairIdx ix, int64_t P, int af, float margin,
                                                            float *__restrict__ pscores,
                                                            float *__restrict__ nscores,
                                                            uint8_t *__restrict__ flags, float *__restrict__ G,
                                                            int32_t *__restrict__ counts,
                                                            int64_t *__restrict__ nviol_accum) {
  constexpr int N = 1 << LOGD;
  extern __shared__ __align__(16) float2 fsm[];
  float2 *bufA = fsm, *bufB = fsm + 3 * N, *tw = fsm + 6 * N;
  float *red = reinterpret_cast<float *>(tw + N / 2);
  for (int m = threadIdx.x; m < N / 2; m += blockDim.x) {
    float sn, cs;
    sincospif(-2.0f * (float)m / (float)N, &sn, &cs);
    tw[m] = make_float2(cs, sn);
  }
  const float inv_n = 1.0f / (float)N;
  int nv = 0;
  for (int64_t i = blockIdx.x; i < P; i += gridDim.x) {
    if (ix.valid && !ix.valid[i]) {
      if (threadIdx.x == 0) {
        flags[i] = 0;
        if (pscores) pscores[i] = 0.f;
        if (nscores) nscores[i] = 0.f;
      }
      continue;
    }
    __syncthreads();
    {
      const float *es = E + (int64_t)ix.sp[i] * N, *eo = E + (int64_t)ix.op[i] * N;
      const float *rp = R + (int64_t)ix.pp[i] * N, *rn = R + (int64_t)ix.pn[i] * N;
      const float *fs = E + (int64_t)ix.sn[i] * N, *fo = E + (int64_t)ix.on[i] * N;
      for (int t = threadIdx.x; t < N; t += blockDim.x) {
        bufA[t] = make_float2(__ldg(es + t), __ldg(eo + t));
        bufA[N + t] = make_float2(__ldg(rp + t), __ldg(rn + t));
        bufA[2 * N + t] = make_float2(__ldg(fs + t), __ldg(fo + t));
      }
    }
    __syncthreads();
    float2 *X = fft_batch<LOGD, 3, false>(bufA, bufB, tw);
    float2 *Y = X == bufA ? bufB : bufA;
    // unpack the spectra, form the products, accumulate the Parseval sums
    float accp = 0.f, accn = 0.f;
    for (int f = threadIdx.x; f < N; f += blockDim.x) {
      const int g = (N - f) & (N - 1);
      const float2 z0 = X[f], z0c = X[g], z1 = X[N + f], z1c = X[N + g], z2 = X[2 * N + f], z2c = X[2 * N + g];
      // x + i y  ->  Xf = (Z_f + conj Z_g) / 2,  Yf = (Z_f - conj Z_g) / (2 i)
      const float2 S = make_float2(0.5f * (z0.x + z0c.x), 0.5f * (z0.y - z0c.y));
      const float2 O = make_float2(0.5f * (z0.y + z0c.y), -0.5f * (z0.x - z0c.x));
      const float2 Rp = make_float2(0.5f * (z1.x + z1c.x), 0.5f * (z1.y - z1c.y));
      const float2 Rn = make_float2(0.5f * (z1.y + z1c.y), -0.5f * (z1.x - z1c.x));
      const float2 S2 = make_float2(0.5f * (z2.x + z2c.x), 0.5f * (z2.y - z2c.y));
      const float2 O2 = make_float2(0.5f * (z2.y + z2c.y), -0.5f * (z2.x - z2c.x));
      const float2 A1 = cmulc(S, O), B1 = cmulc(S2, O2);      // ccorr(s, o)
      const float2 A2 = cmulc(Rp, O), B2 = cmulc(Rn, O2);     // ccorr(r, o)
      const float2 A3 = cmul(S, Rp), B3 = cmul(S2, Rn);       // cconv(s, r)
      accp += A1.x * Rp.x + A1.y * Rp.y;                      // Re(A1 conj(R))
      accn += B1.x * Rn.x + B1.y * Rn.y;
      // two real sequences per inverse transform: U + i V
      Y[f] = make_float2(A1.x - B1.y, A1.y + B1.x);
      Y[N + f] = make_float2(A2.x - A3.y, A2.y + A3.x);
      Y[2 * N + f] = make_float2(B2.x - B3.y, B2.y + B3.x);
    }
    const float raw_p = block_sum(accp, red) * inv_n;
    const float raw_n = block_sum(accn, red) * inv_n;
    const float fp = act_f(af, raw_p), fn = act_f(af, raw_n);
    const bool viol = fn + margin > fp;  // skge/hole.py:56
    if (threadIdx.x == 0) {
      flags[i] = viol;
      if (pscores) pscores[i] = raw_p;
      if (nscores) nscores[i] = raw_n;
    }
    if (!viol) continue;
    ++nv;
    float2 *other = Y == bufA ? bufB : bufA;
    __syncthreads();
    const float2 *Z = fft_batch<LOGD, 3, true>(Y, other, tw);
    const float gp = -act_g_given_f(af, fp) * inv_n, gn = act_g_given_f(af, fn) * inv_n;  // hole.py:66-67
    float *g = G + (int64_t)i * 6 * N;
    const bool same_s = ix.sp[i] == ix.sn[i], same_o = ix.op[i] == ix.on[i], same_r = ix.pp[i] == ix.pn[i];
    for (int t = threadIdx.x; t < N; t += blockDim.x) {
      const float2 y0 = Z[t], y1 = Z[N + t], y2 = Z[2 * N + t];
      // gp ccorr(R[pp], E[op]) -> sp | gn ccorr(R[pn], E[on]) -> sn
      store_folded(g + t, N, same_s, gp * y1.x, gn * y2.x);
      // gp cconv(E[sp], R[pp]) -> op | gn cconv(E[sn], R[pn]) -> on
      store_folded(g + 2 * N + t, N, same_o, gp * y1.y, gn * y2.y);
      // gp ccorr(E[sp], E[op]) -> pp | gn ccorr(E[sn], E[on]) -> pn
      store_folded(g + 4 * N + t, N, same_r, gp * y0.x, gn * y0.y);
    }
  }
  if (threadIdx.x == 0 && nv) {
    atomicAdd(counts, nv);
    if (nviol_accum) atomicAdd(reinterpret_cast<unsigned long long *>(nviol_accum), (unsigned long long)nv);
  }
}

template <int LOGD>
static int launch_hole_fft(const float *E, const float *R, const PairIdx &ix, int64_t P, int af, float margin,
                           float *pscores, float *nscores, uint8_t *flags, float *G, int32_t *counts,
                           int64_t *nviol_accum, cudaStream_t st) {
  constexpr int N = 1 << LOGD;
  size_t smem = (size_t)(6 * N + N / 2) * sizeof(float2) + 40 * sizeof(float);
  SKGE_CUDA(cudaFuncSetAttribute(hole_pair_fft_kernel<LOGD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int threads = N / 4 < 32 ? 32 : (N / 4 > 256 ? 256 : N / 4);  // one radix-4 butterfly index per thread
  int64_t blocks = P > kNumSMs * 32 ? kNumSMs * 32 : P;
  hole_pair_fft_kernel<LOGD><<<(int)blocks, threads, smem, st>>>(E, R, ix, P, af, margin, pscores, nscores, flags, G,
                                                               counts, nviol_accum);
  return 0;
}

// ---------------------------------------------------------------------------
// HolE in the frequency domain (fused training path, power-of-two d).  The trainer keeps
// packed spectra Ehat / Rhat of the parameter tables (csrc/fft.cuh) next to the tables
// themselves, so a pair needs NO transform at all: one warp streams the six spectral rows,
// forms the slot-wise products, reads both scores off by Parseval and, for violating pairs,
// writes the six gradient rows AS SPECTRA.  The segmented reduction sums spectra (the mean is
// linear), and only once per unique row goes back to the time domain, updates the row and
// refreshes its spectrum (segment.cu, spectral mode): 2 transforms per touched row instead
// of 6 per pair.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) hole_spectra_kernel(const float *__restrict__ X, int64_t rows, int d,
                                                           float *__restrict__ Xhat) {
  extern __shared__ __align__(16) float sm_spec[];
  float2 *tw = reinterpret_cast<float2 *>(sm_spec);
  fill_twiddles(tw, d, threadIdx.x, blockDim.x);
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int h = d / 2;
  float2 *b0 = tw + h + (size_t)w * 2 * h, *b1 = b0 + h;
  for (int64_t r = (int64_t)blockIdx.x * nw + w; r < rows; r += (int64_t)gridDim.x * nw) {
    const float2 *x = reinterpret_cast<const float2 *>(X + r * d);
    __syncwarp();
    for (int m = lane; m < h; m += 32) b0[m] = __ldg(x + m);   // z_m = x_{2m} + i x_{2m+1}
    const float2 *Z = warp_rfft_half(b0, b1, tw, d, lane);
    float2 *hr = reinterpret_cast<float2 *>(Xhat + r * d);
    for (int f = lane; f < h; f += 32) hr[f] = packed_slot(Z, f, h, tw);
  }
}

// d = 256: the register-resident transform of fft.cuh (no shared-memory stages; the whole-table
// refresh at the start of an epoch then runs at copy speed)
__global__ void __launch_bounds__(256) hole_spectra256_kernel(const float *__restrict__ X, int64_t rows,
                                                              float *__restrict__ Xhat) {
  __shared__ __align__(16) float4 tbuf[8][64];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  RegFft256 rc;
  regfft256_init(rc, lane);
  for (int64_t r = (int64_t)blockIdx.x * 8 + w; r < rows; r += (int64_t)gridDim.x * 8) {
    const float2 *x = reinterpret_cast<const float2 *>(X + r * 256);
    float2 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = __ldg(x + 32 * j + lane);   // z_m = x_{2m} + i x_{2m+1}, m = 32 j + lane
    regfft256_rfft(v, rc, lane);
    float4 r0, r1;
    regfft256_freq_to_row(tbuf[w], v, r0, r1, lane);
    float4 *hr = reinterpret_cast<float4 *>(Xhat + r * 256);
    hr[lane] = r0;
    hr[32 + lane] = r1;
  }
}

__global__ void __launch_bounds__(256, 4) hole_pair_spec_kernel(const float *__restrict__ Ehat,
                                                             const float *__restrict__ Rhat, PairIdx ix, int64_t P,
                                                             int d, int af, float margin,
                                                             uint8_t *__restrict__ flags, float *__restrict__ G,
                                                             int32_t *__restrict__ counts,
                                                             int64_t *__restrict__ nviol_accum) {
  const int lane = threadIdx.x & 31;
  const int h = d / 2;
  int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const float inv_d = 1.0f / (float)d;
  int nv = 0;
  for (int64_t i = warp; i < P; i += nwarps) {
    if (ix.valid && !ix.valid[i]) {
      if (lane == 0) flags[i] = 0;
      continue;
    }
    const float2 *S = reinterpret_cast<const float2 *>(Ehat + (int64_t)ix.sp[i] * d);
    const float2 *O = reinterpret_cast<const float2 *>(Ehat + (int64_t)ix.op[i] * d);
    const float2 *Rp = reinterpret_cast<const float2 *>(Rhat + (int64_t)ix.pp[i] * d);
    const float2 *S2 = reinterpret_cast<const float2 *>(Ehat + (int64_t)ix.sn[i] * d);
    const float2 *O2 = reinterpret_cast<const float2 *>(Ehat + (int64_t)ix.on[i] * d);
    const float2 *Rn = reinterpret_cast<const float2 *>(Rhat + (int64_t)ix.pn[i] * d);
    float accp = 0.f, accn = 0.f;
    for (int f = lane; f < h; f += 32) {
      const float2 s = __ldg(S + f), o = __ldg(O + f), rp = __ldg(Rp + f);
      const float2 s2 = __ldg(S2 + f), o2 = __ldg(O2 + f), rn = __ldg(Rn + f);
      if (f == 0) {  // slot 0 = (X_0, X_{d/2}), both real
        accp += s.x * o.x * rp.x + s.y * o.y * rp.y;
        accn += s2.x * o2.x * rn.x + s2.y * o2.y * rn.y;
      } else {
        const float2 a1 = cmulc(s, o), b1 = cmulc(s2, o2);
        accp += 2.f * (a1.x * rp.x + a1.y * rp.y);
        accn += 2.f * (b1.x * rn.x + b1.y * rn.y);
      }
    }
    const float raw_p = warp_sum(accp) * inv_d, raw_n = warp_sum(accn) * inv_d;
    const float fp = act_f(af, raw_p), fn = act_f(af, raw_n);
    const bool viol = fn + margin > fp;  // skge/hole.py:56
    if (lane == 0) flags[i] = viol;
    if (!viol) continue;
    ++nv;
    const float gp = -act_g_given_f(af, fp), gn = act_g_given_f(af, fn);  // hole.py:66-67
    float2 *g = reinterpret_cast<float2 *>(G + (int64_t)i * 6 * d);
    const bool same_s = S == S2, same_o = O == O2, same_r = Rp == Rn;
    for (int f = lane; f < h; f += 32) {
      const float2 s = __ldg(S + f), o = __ldg(O + f), rp = __ldg(Rp + f);
      const float2 s2 = __ldg(S2 + f), o2 = __ldg(O2 + f), rn = __ldg(Rn + f);
      float2 a1, a2, a3, b1, b2, b3;
      if (f == 0) {
        a1 = make_float2(s.x * o.x, s.y * o.y);     b1 = make_float2(s2.x * o2.x, s2.y * o2.y);
        a2 = make_float2(rp.x * o.x, rp.y * o.y);   b2 = make_float2(rn.x * o2.x, rn.y * o2.y);
        a3 = make_float2(s.x * rp.x, s.y * rp.y);   b3 = make_float2(s2.x * rn.x, s2.y * rn.y);
      } else {
        a1 = cmulc(s, o);   b1 = cmulc(s2, o2);     // ccorr(s, o)
        a2 = cmulc(rp, o);  b2 = cmulc(rn, o2);     // ccorr(r, o)
        a3 = cmul(s, rp);   b3 = cmul(s2, rn);      // cconv(s, r)
      }
      store_folded(g + 0 * h + f, h, same_s, make_float2(gp * a2.x, gp * a2.y), make_float2(gn * b2.x, gn * b2.y));  // -> sp | sn
      store_folded(g + 2 * h + f, h, same_o, make_float2(gp * a3.x, gp * a3.y), make_float2(gn * b3.x, gn * b3.y));  // -> op | on
      store_folded(g + 4 * h + f, h, same_r, make_float2(gp * a1.x, gp * a1.y), make_float2(gn * b1.x, gn * b1.y));  // -> pp | pn
    }
  }
  if (lane == 0 && nv) {
    atomicAdd(counts, nv);
    if (nviol_accum) atomicAdd(reinterpret_cast<unsigned long long *>(nviol_accum), (unsigned long long)nv);
  }
}

// The same kernel for d = 128 * NIT with 128-bit accesses: a lane owns the complex slots
// (2 lane + 64 it, 2 lane + 64 it + 1), the six spectral rows stay in registers between the score
// and the gradient rows, so every row is requested once and each request / store moves 16 bytes.
//
// Relation rows are PRE-REDUCED: the pairs are visited in relation order (`order`, a stable sort of
// the pair indices by pp), a warp walks KBLK consecutive pairs of that order and keeps the sum of
// their relation-gradient spectra in registers while the relation does not change; one row per run
// is written (into the relation row of the run's first violating pair, `runw` = occurrences it
// stands for, 0 for the other pairs of the run).  With M relations this turns P relation rows
// (a quarter of G, and the hot-row chunk pass that had to read them) into about P / KBLK + M.
#ifndef SKGE_PAIR_SPEC4_CTAS
#define SKGE_PAIR_SPEC4_CTAS 3
#endif
// state of the relation run a warp is summing
template <int NIT>
struct RelRun {
  int p = -1, cnt = 0;
  int64_t head = 0;
  float4 acc[NIT];
};

template <int NIT>
__device__ __forceinline__ void rel_run_flush(RelRun<NIT> &run, float *__restrict__ G, int32_t *__restrict__ runw, int lane) {
  if (run.p < 0) return;
  float4 *gr = reinterpret_cast<float4 *>(G + (run.head * 6 + 4) * (128 * NIT));
#pragma unroll
  for (int it = 0; it < NIT; ++it) gr[lane + 32 * it] = run.acc[it];
  if (lane == 0) runw[run.head] = run.cnt;
  run.p = -1;
}

// slot-wise products of packed spectra; slot 0 of a row holds two REAL spectral values
__device__ __forceinline__ float2 spec_mul(float2 a, float2 b, bool real0) {      // a b
  return real0 ? make_float2(a.x * b.x, a.y * b.y) : cmul(a, b);
}
__device__ __forceinline__ float2 spec_mulc(float2 a, float2 b, bool real0) {     // conj(a) b
  return real0 ? make_float2(a.x * b.x, a.y * b.y) : cmulc(a, b);
}

// One pair of any shape (both or no entity slot shared, a corrupted relation, ...): six rows through
// registers, relation rows written per pair (runw = 2 for a folded row, 1 otherwise).  Rare on the
// fused path: the staged kernel only lists such pairs, a second small kernel works the list off, so
// that their register needs do not shape the staged kernel.
template <int NIT>
__device__ __forceinline__ int hole_spec4_generic_pair(const float *__restrict__ Ehat, const float *__restrict__ Rhat,
                                                    const PairIdx &ix, int64_t i, int af, float margin,
                                                    uint8_t *__restrict__ flags, float *__restrict__ G,
                                                    int32_t *__restrict__ runw, int lane) {
  constexpr int d = 128 * NIT, h = d / 2;
  const float inv_d = 1.0f / (float)d;
  const float4 *S = reinterpret_cast<const float4 *>(Ehat + (int64_t)ix.sp[i] * d);
  const float4 *O = reinterpret_cast<const float4 *>(Ehat + (int64_t)ix.op[i] * d);
  const float4 *Rp = reinterpret_cast<const float4 *>(Rhat + (int64_t)ix.pp[i] * d);
  const float4 *S2 = reinterpret_cast<const float4 *>(Ehat + (int64_t)ix.sn[i] * d);
  const float4 *O2 = reinterpret_cast<const float4 *>(Ehat + (int64_t)ix.on[i] * d);
  const float4 *Rn = reinterpret_cast<const float4 *>(Rhat + (int64_t)ix.pn[i] * d);
  float4 s[NIT], o[NIT], rp[NIT], s2[NIT], o2[NIT], rn[NIT];
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int f4 = lane + 32 * it;
    s[it] = __ldg(S + f4); o[it] = __ldg(O + f4); rp[it] = __ldg(Rp + f4);
    s2[it] = __ldg(S2 + f4); o2[it] = __ldg(O2 + f4); rn[it] = __ldg(Rn + f4);
  }
  float accp = 0.f, accn = 0.f;
#pragma unroll
  for (int it = 0; it < NIT; ++it)
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const bool real0 = it == 0 && u == 0 && lane == 0;
      const float2 sv = u ? make_float2(s[it].z, s[it].w) : make_float2(s[it].x, s[it].y);
      const float2 ov = u ? make_float2(o[it].z, o[it].w) : make_float2(o[it].x, o[it].y);
      const float2 rv = u ? make_float2(rp[it].z, rp[it].w) : make_float2(rp[it].x, rp[it].y);
      const float2 sv2 = u ? make_float2(s2[it].z, s2[it].w) : make_float2(s2[it].x, s2[it].y);
      const float2 ov2 = u ? make_float2(o2[it].z, o2[it].w) : make_float2(o2[it].x, o2[it].y);
      const float2 rv2 = u ? make_float2(rn[it].z, rn[it].w) : make_float2(rn[it].x, rn[it].y);
      const float2 a1 = spec_mulc(sv, ov, real0), b1 = spec_mulc(sv2, ov2, real0);
      const float wgt = real0 ? 1.f : 2.f;
      accp += wgt * (a1.x * rv.x + a1.y * rv.y);
      accn += wgt * (b1.x * rv2.x + b1.y * rv2.y);
    }
  const float raw_p = warp_sum(accp) * inv_d, raw_n = warp_sum(accn) * inv_d;
  const float fp = act_f(af, raw_p), fn = act_f(af, raw_n);
  const bool viol = fn + margin > fp;  // skge/hole.py:56
  if (lane == 0) flags[i] = viol;
  if (!viol) return 0;
  const float gp = -act_g_given_f(af, fp), gn = act_g_given_f(af, fn);  // hole.py:66-67
  float4 *g = reinterpret_cast<float4 *>(G + (int64_t)i * 6 * d);
  const bool same_s = S == S2, same_o = O == O2, same_r = Rp == Rn;
  if (lane == 0) runw[i] = same_r ? 2 : 1;
  constexpr int h4 = h / 2;   // float4 per spectral row
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int f4 = lane + 32 * it;
    float2 a1[2], a2[2], a3[2], b1[2], b2[2], b3[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const bool real0 = it == 0 && u == 0 && lane == 0;
      const float2 sv = u ? make_float2(s[it].z, s[it].w) : make_float2(s[it].x, s[it].y);
      const float2 ov = u ? make_float2(o[it].z, o[it].w) : make_float2(o[it].x, o[it].y);
      const float2 rv = u ? make_float2(rp[it].z, rp[it].w) : make_float2(rp[it].x, rp[it].y);
      const float2 sv2 = u ? make_float2(s2[it].z, s2[it].w) : make_float2(s2[it].x, s2[it].y);
      const float2 ov2 = u ? make_float2(o2[it].z, o2[it].w) : make_float2(o2[it].x, o2[it].y);
      const float2 rv2 = u ? make_float2(rn[it].z, rn[it].w) : make_float2(rn[it].x, rn[it].y);
      a1[u] = spec_mulc(sv, ov, real0);  b1[u] = spec_mulc(sv2, ov2, real0);   // ccorr(s, o)
      a2[u] = spec_mulc(rv, ov, real0);  b2[u] = spec_mulc(rv2, ov2, real0);   // ccorr(r, o)
      a3[u] = spec_mul(sv, rv, real0);   b3[u] = spec_mul(sv2, rv2, real0);    // cconv(s, r)
    }
    auto put = [&](int row, bool same, const float2 (&x)[2], const float2 (&y)[2]) {
      const float4 vp = make_float4(gp * x[0].x, gp * x[0].y, gp * x[1].x, gp * x[1].y);
      const float4 vn = make_float4(gn * y[0].x, gn * y[0].y, gn * y[1].x, gn * y[1].y);
      float4 *dst = g + row * h4 + f4;
      if (same) {
        dst[0] = make_float4(vp.x + vn.x, vp.y + vn.y, vp.z + vn.z, vp.w + vn.w);
      } else {
        dst[0] = vp;
        dst[h4] = vn;
      }
    };
    put(0, same_s, a2, b2);  // -> sp | sn
    put(2, same_o, a3, b3);  // -> op | on
    put(4, same_r, a1, b1);  // -> pp | pn
  }
  return 1;
}

// A corrupted pair whose rows are staged in shared memory: S = E^[sp], O = E^[op], C = the
// corrupted entity's row, R = R^[pp].  KIND 1: the object is corrupted (sn = sp, on = c),
// KIND 2: the subject is (sn = c, on = op); pn = pp in both.  Four rows in, three entity gradient
// rows out (the shared slot's row already folded), the relation gradient joins the warp's run.
template <int NIT, int KIND>
__device__ __forceinline__ int hole_spec4_staged_pair(const float4 *S, const float4 *O, const float4 *C,
                                                      const float4 *R, int64_t i, int prel, int af, float margin,
                                                      uint8_t *__restrict__ flags, float *__restrict__ G,
                                                      int32_t *__restrict__ runw, float *__restrict__ coef,
                                                      RelRun<NIT> &run, int lane) {
  constexpr int d = 128 * NIT, h4 = d / 4;
  const float inv_d = 1.0f / (float)d;
  // (the rows are read from the stage twice, for the scores and for the gradient rows: one 128-bit
  //  slice of each row is live at a time instead of the whole row)
  float accp = 0.f, accn = 0.f;
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int f4 = lane + 32 * it;
    const float4 s4 = S[f4], o4 = O[f4], c4 = C[f4], r4 = R[f4];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const bool real0 = it == 0 && u == 0 && lane == 0;
      const float2 sv = u ? make_float2(s4.z, s4.w) : make_float2(s4.x, s4.y);
      const float2 ov = u ? make_float2(o4.z, o4.w) : make_float2(o4.x, o4.y);
      const float2 cv = u ? make_float2(c4.z, c4.w) : make_float2(c4.x, c4.y);
      const float2 rv = u ? make_float2(r4.z, r4.w) : make_float2(r4.x, r4.y);
      const float2 a1 = spec_mulc(sv, ov, real0);
      const float2 b1 = KIND == 1 ? spec_mulc(sv, cv, real0) : spec_mulc(cv, ov, real0);
      const float wgt = real0 ? 1.f : 2.f;
      accp += wgt * (a1.x * rv.x + a1.y * rv.y);
      accn += wgt * (b1.x * rv.x + b1.y * rv.y);
    }
  }
  const float raw_p = warp_sum(accp) * inv_d, raw_n = warp_sum(accn) * inv_d;
  const float fp = act_f(af, raw_p), fn = act_f(af, raw_n);
  const bool viol = fn + margin > fp;  // skge/hole.py:56
  if (lane == 0) flags[i] = viol;
  if (!viol) return 0;
  const float gp = -act_g_given_f(af, fp), gn = act_g_given_f(af, fn);  // hole.py:66-67
  if (run.p != prel) {
    rel_run_flush<NIT>(run, G, runw, lane);
    run.p = prel; run.head = i; run.cnt = 0;
#pragma unroll
    for (int it = 0; it < NIT; ++it) run.acc[it] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  run.cnt += 2;   // the positive's and the negative's occurrence of the relation row
  if (lane == 0) {
    if (i != run.head) runw[i] = 0;
    *reinterpret_cast<float2 *>(coef + 2 * i) = make_float2(gp, gn);
  }
  float4 *g = reinterpret_cast<float4 *>(G + (int64_t)i * 6 * d);
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int f4 = lane + 32 * it;
    const float4 s4 = S[f4], o4 = O[f4], c4 = C[f4], r4 = R[f4];
    float2 x0[2], x2[2], rel[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const bool real0 = it == 0 && u == 0 && lane == 0;
      const float2 sv = u ? make_float2(s4.z, s4.w) : make_float2(s4.x, s4.y);
      const float2 ov = u ? make_float2(o4.z, o4.w) : make_float2(o4.x, o4.y);
      const float2 cv = u ? make_float2(c4.z, c4.w) : make_float2(c4.x, c4.y);
      const float2 rv = u ? make_float2(r4.z, r4.w) : make_float2(r4.x, r4.y);
      if (KIND == 1) {
        const float2 wv = make_float2(gp * ov.x + gn * cv.x, gp * ov.y + gn * cv.y);
        x0[u] = spec_mulc(rv, wv, real0);     // row 0 -> sp (= sn): ccorr(r, gp o + gn c)
        x2[u] = spec_mul(sv, rv, real0);      // row 2 = cconv(s, r): -> op times gp, -> on times gn (RoleMap::coef)
        rel[u] = spec_mulc(sv, wv, real0);    // -> pp: ccorr(s, gp o + gn c)
      } else {
        const float2 wv = make_float2(gp * sv.x + gn * cv.x, gp * sv.y + gn * cv.y);
        x0[u] = spec_mulc(rv, ov, real0);     // row 0 = ccorr(r, o): -> sp times gp, -> sn times gn
        x2[u] = spec_mul(wv, rv, real0);      // row 2 -> op (= on): cconv(gp s + gn c, r)
        rel[u] = spec_mulc(wv, ov, real0);    // -> pp: ccorr(gp s + gn c, o)
      }
    }
    g[f4] = make_float4(x0[0].x, x0[0].y, x0[1].x, x0[1].y);
    g[2 * h4 + f4] = make_float4(x2[0].x, x2[0].y, x2[1].x, x2[1].y);
    run.acc[it].x += rel[0].x; run.acc[it].y += rel[0].y; run.acc[it].z += rel[1].x; run.acc[it].w += rel[1].y;
  }
  return 1;
}

__host__ __device__ constexpr size_t pair_spec4_smem_bytes(int nit) { return 128 + (size_t)8 * 2 * 4 * 512 * nit; }

// Driver: a warp takes KBLK = 32 consecutive pairs of the relation order, one per lane for the index
// work.  The four rows of a corrupted pair (kinds 1 and 2) come in by bulk-TMA copies, double
// buffered per warp: the copies of the next pair are in flight while this one is computed.  Pairs of
// any other shape (both or no entity slot shared, a corrupted relation) take the six-row register
// path above afterwards.
template <int NIT>
__global__ void __launch_bounds__(256, SKGE_PAIR_SPEC4_CTAS) hole_pair_spec4_kernel(
    const float *__restrict__ Ehat, const float *__restrict__ Rhat, PairIdx ix, int64_t P, int af, float margin,
    uint8_t *__restrict__ flags, float *__restrict__ G, int32_t *__restrict__ counts, int64_t *__restrict__ nviol_accum,
    const int32_t *__restrict__ order, int32_t *__restrict__ runw, float *__restrict__ coef,
    int32_t *__restrict__ glist) {
  extern __shared__ __align__(128) unsigned char pair_smem[];
  constexpr int KBLK = 32, d = 128 * NIT, ROW4 = d / 4;
  constexpr uint32_t ROWB = d * 4;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint64_t *bars = reinterpret_cast<uint64_t *>(pair_smem) + 2 * w;
  float4 *buf = reinterpret_cast<float4 *>(pair_smem + 128) + (size_t)w * 2 * 4 * ROW4;   // [stage][slot][ROW4]
  if (lane == 0) {
    ptx::mbar_init(bars, 1);
    ptx::mbar_init(bars + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + w;
  int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t nblk = (P + KBLK - 1) / KBLK;
  uint32_t issued = 0, done = 0;
  int nv = 0;
  for (int64_t blk = warp; blk < nblk; blk += nwarps) {
    // lane l: the indices of pair l of the block
    const int64_t pos = blk * KBLK + lane;
    int my_i = -1, kind = 0, rowa = 0, rowb = 0, rowc = 0, rowr = 0;
    if (pos < P) {
      my_i = order[pos];
      if (ix.valid && !ix.valid[my_i]) {
        flags[my_i] = 0;
      } else {
        const int sp = ix.sp[my_i], op = ix.op[my_i], pp = ix.pp[my_i];
        const int sn = ix.sn[my_i], on = ix.on[my_i], pn = ix.pn[my_i];
        rowa = sp; rowb = op; rowr = pp;
        if (pp == pn && sp == sn && op != on) { kind = 1; rowc = on; }
        else if (pp == pn && op == on && sp != sn) { kind = 2; rowc = sn; }
        else kind = 3;
      }
    }
    RelRun<NIT> run;
    auto issue = [&](int t) {
      const int st = issued & 1;
      const int ra = __shfl_sync(kFull, rowa, t), rb = __shfl_sync(kFull, rowb, t);
      const int rcc = __shfl_sync(kFull, rowc, t), rr = __shfl_sync(kFull, rowr, t);
      // the stage was last read by this warp's own (generic-proxy) loads
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) ptx::mbar_expect_tx(bars + st, 4 * ROWB);
      __syncwarp();
      if (lane < 4) {
        const float *src = lane == 3 ? Rhat + (int64_t)rr * d : Ehat + (int64_t)(lane == 0 ? ra : lane == 1 ? rb : rcc) * d;
        ptx::bulk_g2s(buf + (st * 4 + lane) * ROW4, src, ROWB, bars + st);
      }
      ++issued;
    };
    unsigned m = __ballot_sync(kFull, kind == 1 || kind == 2);
    if (m) issue(__ffs(m) - 1);
    while (m) {
      const int t = __ffs(m) - 1;
      m &= m - 1;
      if (m) issue(__ffs(m) - 1);
      const int st = done & 1;
      ptx::mbar_wait(bars + st, (done >> 1) & 1);
      ++done;
      const int64_t i = __shfl_sync(kFull, my_i, t);
      const int prel = __shfl_sync(kFull, rowr, t);
      const float4 *sl = buf + st * 4 * ROW4;
      if (__shfl_sync(kFull, kind, t) == 1)
        nv += hole_spec4_staged_pair<NIT, 1>(sl, sl + ROW4, sl + 2 * ROW4, sl + 3 * ROW4, i, prel, af, margin, flags, G, runw, coef, run, lane);
      else
        nv += hole_spec4_staged_pair<NIT, 2>(sl, sl + ROW4, sl + 2 * ROW4, sl + 3 * ROW4, i, prel, af, margin, flags, G, runw, coef, run, lane);
    }
    rel_run_flush<NIT>(run, G, runw, lane);
    if (kind == 3) glist[atomicAdd(counts + 3, 1)] = my_i;   // any order: these pairs are independent of each other
  }
  if (lane == 0 && nv) {
    atomicAdd(counts, nv);
    if (nviol_accum) atomicAdd(reinterpret_cast<unsigned long long *>(nviol_accum), (unsigned long long)nv);
  }
}

// the pairs the staged kernel listed in glist[0 .. counts[3])
template <int NIT>
__global__ void __launch_bounds__(256) hole_pair_spec4_generic_kernel(
    const float *__restrict__ Ehat, const float *__restrict__ Rhat, PairIdx ix, int af, float margin,
    uint8_t *__restrict__ flags, float *__restrict__ G, int32_t *__restrict__ counts, int64_t *__restrict__ nviol_accum,
    int32_t *__restrict__ runw, const int32_t *__restrict__ glist) {
  const int lane = threadIdx.x & 31;
  const int n = counts[3];
  int nv = 0;
  for (int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); k < n; k += gridDim.x * (blockDim.x >> 5))
    nv += hole_spec4_generic_pair<NIT>(Ehat, Rhat, ix, glist[k], af, margin, flags, G, runw, lane);
  if (lane == 0 && nv) {
    atomicAdd(counts, nv);
    if (nviol_accum) atomicAdd(reinterpret_cast<unsigned long long *>(nviol_accum), (unsigned long long)nv);
  }
}

static int pair_block_threads(int d) {  // four groups of d4/4 threads (hole_pair_kernel)
  int t = (round4(d) + 31) / 32 * 32;
  return t < 64 ? 64 : t;
}

struct PairBuffers {
  uint8_t *flags;
  float *G;
};

static size_t pair_ws_bytes(int64_t P, int d, int rows, int nroles) {
  if (P < 1) P = 1;
  return align_up((size_t)P) + align_up((size_t)P * rows * d * sizeof(float)) +
         seg_workspace_bytes((int64_t)nroles * P, d) + order_workspace_bytes(P) + align_up((size_t)P * 4) +
         align_up((size_t)P * 8) + align_up((size_t)P * 4) + 1024;
}

// model: 0 TransE, 1 HolE
static int pair_run(int model, float *E, float *R, float *p2E, float *p2R, const PairIdx &ix, int64_t P,
                    int64_t N, int64_t M, int d, int l1_or_af, float margin, float rparam, bool update,
                    int opt, float lr, int postE, int postR, float *pscores, float *nscores, float *ge,
                    int32_t *eidx, float *gr, int32_t *ridx, int32_t *counts, int64_t *nviol_accum,
                    int32_t *ent_viol, int32_t *ucE, int32_t *ucR, void *ws, size_t ws_bytes,
                    cudaStream_t st, float *Ehat = nullptr, float *Rhat = nullptr) {
  SKGE_REQUIRE(E && R && ix.sp && ix.op && ix.pp && ix.sn && ix.on && ix.pn && counts && ws,
               "null argument");
  SKGE_REQUIRE(P > 0 && d > 0 && N > 0 && M > 0, "bad sizes");
  if (update) SKGE_REQUIRE(opt == SKGE_OPT_SGD || (p2E && p2R), "AdaGrad needs p2E/p2R");
  else SKGE_REQUIRE(ge && eidx && gr && ridx, "null output");
  const int rows = model == 0 ? 2 : 6;
  Arena ar(ws, ws_bytes);
  uint8_t *flags = ar.take<uint8_t>(P);
  float *G = ar.take<float>((size_t)P * rows * d);
  if (!ar.ok()) {
    set_error("workspace too small: need > %zu bytes, have %zu", ar.off, ar.cap);
    return SKGE_EWORKSPACE;
  }
  SKGE_CUDA(cudaMemsetAsync(counts, 0, 4 * sizeof(int32_t), st));
  int32_t *runw = nullptr;
  float *coef = nullptr;
  if (model == 0) {
    int64_t blocks = (P + 7) / 8;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    const int vec = pick_vec(d);
    const int chunks = (d + 32 * vec - 1) / (32 * vec);
#define SKGE_TRANSE_PAIR(KERNEL) \
    KERNEL<<<(int)blocks, 256, 0, st>>>(E, R, ix, P, d, l1_or_af, margin, pscores, nscores, flags, G, counts, \
                                        nviol_accum, ent_viol)
    if (vec == 4 && chunks == 1) SKGE_TRANSE_PAIR((transe_pair_reg_kernel<4, 1>));
    else if (vec == 4 && chunks == 2) SKGE_TRANSE_PAIR((transe_pair_reg_kernel<4, 2>));
    else if (vec == 2 && chunks == 1) SKGE_TRANSE_PAIR((transe_pair_reg_kernel<2, 1>));
    else if (vec == 2 && chunks == 2) SKGE_TRANSE_PAIR((transe_pair_reg_kernel<2, 2>));
    else if (vec == 1 && chunks == 1) SKGE_TRANSE_PAIR((transe_pair_reg_kernel<1, 1>));
    else if (vec == 1 && chunks == 2) SKGE_TRANSE_PAIR((transe_pair_reg_kernel<1, 2>));
    else if (vec == 4) SKGE_TRANSE_PAIR(transe_pair_kernel<4>);
    else if (vec == 2) SKGE_TRANSE_PAIR(transe_pair_kernel<2>);
    else SKGE_TRANSE_PAIR(transe_pair_kernel<1>);
#undef SKGE_TRANSE_PAIR
  } else if (Ehat && Rhat) {
    SKGE_REQUIRE(update && spectral_len_ok(d), "spectral HolE step needs an even d in [32, 1024] with d / 2 = 2^a 3^b 5^c");
    int64_t blocks = (P + 7) / 8;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    if (d == 128 || d == 256) {
      // 128-bit accesses, register-resident rows, relation rows pre-reduced over relation-ordered pairs
      const int32_t *order = nullptr;
      int kb = 1;
      while (((int64_t)1 << kb) < M) ++kb;
      if (int rc = order_by_key(ix.pp, P, kb, ar, st, &order)) return rc;
      runw = ar.take<int32_t>(P);
      coef = ar.take<float>(2 * P);
      int32_t *glist = ar.take<int32_t>(P);
      if (!ar.ok()) {
        set_error("workspace too small: need > %zu bytes, have %zu", ar.off, ar.cap);
        return SKGE_EWORKSPACE;
      }
      SKGE_REQUIRE(((reinterpret_cast<uintptr_t>(Ehat) | reinterpret_cast<uintptr_t>(Rhat)) & 15) == 0,
                   "spectral tables must be 16-byte aligned");
      blocks = ((P + 31) / 32 + 7) / 8;
      if (blocks > kNumSMs * SKGE_PAIR_SPEC4_CTAS) blocks = kNumSMs * SKGE_PAIR_SPEC4_CTAS;
      if (d == 128) {
        const size_t smem = pair_spec4_smem_bytes(1);
        SKGE_CUDA(cudaFuncSetAttribute(hole_pair_spec4_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        hole_pair_spec4_kernel<1><<<(int)blocks, 256, smem, st>>>(Ehat, Rhat, ix, P, l1_or_af, margin, flags, G, counts, nviol_accum, order, runw, coef, glist);
        hole_pair_spec4_generic_kernel<1><<<kNumSMs, 256, 0, st>>>(Ehat, Rhat, ix, l1_or_af, margin, flags, G, counts, nviol_accum, runw, glist);
      } else {
        const size_t smem = pair_spec4_smem_bytes(2);
        SKGE_CUDA(cudaFuncSetAttribute(hole_pair_spec4_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        hole_pair_spec4_kernel<2><<<(int)blocks, 256, smem, st>>>(Ehat, Rhat, ix, P, l1_or_af, margin, flags, G, counts, nviol_accum, order, runw, coef, glist);
        hole_pair_spec4_generic_kernel<2><<<kNumSMs, 256, 0, st>>>(Ehat, Rhat, ix, l1_or_af, margin, flags, G, counts, nviol_accum, runw, glist);
      }
    } else {
      hole_pair_spec_kernel<<<(int)blocks, 256, 0, st>>>(Ehat, Rhat, ix, P, d, l1_or_af, margin, flags, G, counts,
                                                        nviol_accum);
    }
  } else {
    SKGE_REQUIRE(d <= 1024, "HolE pair kernel supports d <= 1024");
    bool fft_done = true;
    switch (d) {  // power-of-two d: shared-memory FFT path
      case 32: if (int rc = launch_hole_fft<5>(E, R, ix, P, l1_or_af, margin, pscores, nscores, flags, G, counts, nviol_accum, st)) return rc; break;
      case 64: if (int rc = launch_hole_fft<6>(E, R, ix, P, l1_or_af, margin, pscores, nscores, flags, G, counts, nviol_accum, st)) return rc; break;
      case 128: if (int rc = launch_hole_fft<7>(E, R, ix, P, l1_or_af, margin, pscores, nscores, flags, G, counts, nviol_accum, st)) return rc; break;
      case 256: if (int rc = launch_hole_fft<8>(E, R, ix, P, l1_or_af, margin, pscores, nscores, flags, G, counts, nviol_accum, st)) return rc; break;
      case 512: if (int rc = launch_hole_fft<9>(E, R, ix, P, l1_or_af, margin, pscores, nscores, flags, G, counts, nviol_accum, st)) return rc; break;
      case 1024: if (int rc = launch_hole_fft<10>(E, R, ix, P, l1_or_af, margin, pscores, nscores, flags, G, counts, nviol_accum, st)) return rc; break;
      default: fft_done = false;
    }
    if (!fft_done) {
    size_t smem = (20 * (size_t)round4(d) + 40) * sizeof(float);
    SKGE_CUDA(cudaFuncSetAttribute(hole_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = P > kNumSMs * 16 ? kNumSMs * 16 : P;
    hole_pair_kernel<<<(int)blocks, pair_block_threads(d), smem, st>>>(E, R, ix, P, d, l1_or_af, margin,
                                                                      pscores, nscores, flags, G, counts,
                                                                      nviol_accum);
    }
  }
  SKGE_LAUNCH_CHECK();

  RoleMap rm;
  if (model == 0) {
    // entity keys sp+op+sn+on get (pg,-pg,ng,-ng); relation keys pp+pn get (pg,ng): transe.py:128-160
    const int32_t *idx[6] = {ix.sp, ix.op, ix.sn, ix.on, ix.pp, ix.pn};
    const int isrel[6] = {0, 0, 0, 0, 1, 1}, grow[6] = {0, 0, 1, 1, 0, 1};
    const float sgn[6] = {1.f, -1.f, 1.f, -1.f, 1.f, 1.f};
    for (int r = 0; r < 6; ++r) { rm.idx[r] = idx[r]; rm.is_rel[r] = isrel[r]; rm.grow[r] = grow[r]; rm.gsign[r] = sgn[r]; }
  } else {
    // entity keys sp+sn+op+on, relation keys pp+pn: hole.py:69-97
    const int32_t *idx[6] = {ix.sp, ix.sn, ix.op, ix.on, ix.pp, ix.pn};
    const int isrel[6] = {0, 0, 0, 0, 1, 1}, grow[6] = {0, 1, 2, 3, 4, 5};
    for (int r = 0; r < 6; ++r) {
      rm.idx[r] = idx[r]; rm.is_rel[r] = isrel[r]; rm.grow[r] = grow[r]; rm.gsign[r] = 1.f;
      rm.twin[r] = r ^ 1;  // the kernels above fold rows (2q, 2q + 1) when the ids coincide
    }
    if (runw) { rm.runw = runw; rm.runw_role = 4; rm.coef = coef; }
  }
  rm.nroles = 6;
  ParamDesc pd[2];
  pd[0] = ParamDesc{E, p2E, postE, 0.f, ucE, ge, eidx, Ehat};
  pd[1] = ParamDesc{R, p2R, postR, rparam, ucR, gr, ridx, Rhat};
  return seg_run(rm, flags, P, N, M, d, G, rows, pd, update, opt, lr, counts, ar, st,
                 (model == 1 && Ehat && Rhat) ? 1 : 0);
}

}  // namespace skge

using namespace skge;

extern "C" {

size_t skge_pair_workspace_bytes(int64_t P, int d, int rows_per_pair, int64_t N, int64_t M) {
  (void)N; (void)M;
  return pair_ws_bytes(P, d, rows_per_pair, 6);
}

int skge_transe_pair_grads(const float *E, const float *R, const int32_t *sp, const int32_t *op,
                           const int32_t *pp, const int32_t *sn, const int32_t *on,
                           const int32_t *pn, const uint8_t *valid, int64_t P, int64_t N,
                           int64_t M, int d, int l1, float margin, float *pscores,
                           float *nscores, float *ge, int32_t *eidx, float *gr, int32_t *ridx,
                           int32_t *counts, int32_t *ent_violations, void *ws, size_t ws_bytes,
                           skge_stream_t stream) {
  PairIdx ix{sp, op, pp, sn, on, pn, valid};
  return pair_run(0, const_cast<float *>(E), const_cast<float *>(R), nullptr, nullptr, ix, P, N, M, d, l1,
                  margin, 0.f, false, SKGE_OPT_SGD, 0.f, SKGE_POST_NONE, SKGE_POST_NONE, pscores, nscores, ge,
                  eidx, gr, ridx, counts, nullptr, ent_violations, nullptr, nullptr, ws, ws_bytes,
                  as_stream(stream));
}

int skge_transe_pair_step(float *E, float *R, float *p2E, float *p2R, const int32_t *sp,
                          const int32_t *op, const int32_t *pp, const int32_t *sn,
                          const int32_t *on, const int32_t *pn, const uint8_t *valid, int64_t P,
                          int64_t N, int64_t M, int d, int l1, float margin, int opt, float lr,
                          int postE, int postR, int32_t *counts, int64_t *nviol_accum,
                          int32_t *ent_violations, int32_t *upd_counts_E, int32_t *upd_counts_R,
                          void *ws, size_t ws_bytes, skge_stream_t stream) {
  PairIdx ix{sp, op, pp, sn, on, pn, valid};
  return pair_run(0, E, R, p2E, p2R, ix, P, N, M, d, l1, margin, 0.f, true, opt, lr, postE, postR, nullptr,
                  nullptr, nullptr, nullptr, nullptr, nullptr, counts, nviol_accum, ent_violations,
                  upd_counts_E, upd_counts_R, ws, ws_bytes, as_stream(stream));
}

int skge_hole_pair_grads(const float *E, const float *R, const int32_t *sp, const int32_t *op,
                         const int32_t *pp, const int32_t *sn, const int32_t *on,
                         const int32_t *pn, const uint8_t *valid, int64_t P, int64_t N, int64_t M,
                         int d, int af, float margin, float rparam, float *pscores,
                         float *nscores, float *ge, int32_t *eidx, float *gr, int32_t *ridx,
                         int32_t *counts, void *ws, size_t ws_bytes, skge_stream_t stream) {
  PairIdx ix{sp, op, pp, sn, on, pn, valid};
  return pair_run(1, const_cast<float *>(E), const_cast<float *>(R), nullptr, nullptr, ix, P, N, M, d, af,
                  margin, rparam, false, SKGE_OPT_SGD, 0.f, SKGE_POST_NONE, SKGE_POST_NONE, pscores, nscores,
                  ge, eidx, gr, ridx, counts, nullptr, nullptr, nullptr, nullptr, ws, ws_bytes,
                  as_stream(stream));
}

int skge_hole_pair_step(float *E, float *R, float *p2E, float *p2R, const int32_t *sp,
                        const int32_t *op, const int32_t *pp, const int32_t *sn,
                        const int32_t *on, const int32_t *pn, const uint8_t *valid, int64_t P,
                        int64_t N, int64_t M, int d, int af, float margin, float rparam, int opt,
                        float lr, int postE, int postR, int32_t *counts, int64_t *nviol_accum,
                        int32_t *upd_counts_E, int32_t *upd_counts_R, void *ws, size_t ws_bytes,
                        skge_stream_t stream) {
  PairIdx ix{sp, op, pp, sn, on, pn, valid};
  return pair_run(1, E, R, p2E, p2R, ix, P, N, M, d, af, margin, rparam, true, opt, lr, postE, postR, nullptr,
                  nullptr, nullptr, nullptr, nullptr, nullptr, counts, nviol_accum, nullptr, upd_counts_E,
                  upd_counts_R, ws, ws_bytes, as_stream(stream));
}


int skge_hole_spectra(const float *X, int64_t rows, int d, float *Xhat, skge_stream_t stream) {
  SKGE_REQUIRE(X && Xhat && rows >= 0, "bad arguments");
  SKGE_REQUIRE(spectral_len_ok(d), "spectra need an even d in [32, 1024] with d / 2 = 2^a 3^b 5^c");
  if (rows == 0) return 0;
  if (d == 256 && ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Xhat)) & 15) == 0) {
    int64_t nb = (rows + 7) / 8;
    if (nb > kNumSMs * 8) nb = kNumSMs * 8;
    hole_spectra256_kernel<<<(int)nb, 256, 0, as_stream(stream)>>>(X, rows, Xhat);
    SKGE_LAUNCH_CHECK();
    return 0;
  }
  size_t smem = (size_t)(d / 2) * sizeof(float2) + (size_t)8 * warp_fft_scratch_floats(d) * sizeof(float);
  SKGE_CUDA(cudaFuncSetAttribute(hole_spectra_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t blocks = (rows + 7) / 8;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  hole_spectra_kernel<<<(int)blocks, 256, smem, as_stream(stream)>>>(X, rows, d, Xhat);
  SKGE_LAUNCH_CHECK();
  return 0;
}

int skge_hole_pair_step_spectral(float *E, float *R, float *Ehat, float *Rhat, float *p2E, float *p2R,
                                 const int32_t *sp, const int32_t *op, const int32_t *pp, const int32_t *sn,
                                 const int32_t *on, const int32_t *pn, const uint8_t *valid, int64_t P,
                                 int64_t N, int64_t M, int d, int af, float margin, float rparam, int opt,
                                 float lr, int postE, int postR, int32_t *counts, int64_t *nviol_accum,
                                 int32_t *upd_counts_E, int32_t *upd_counts_R, void *ws, size_t ws_bytes,
                                 skge_stream_t stream) {
  SKGE_REQUIRE(Ehat && Rhat, "null spectra");
  PairIdx ix{sp, op, pp, sn, on, pn, valid};
  return pair_run(1, E, R, p2E, p2R, ix, P, N, M, d, af, margin, rparam, true, opt, lr, postE, postR, nullptr,
                  nullptr, nullptr, nullptr, nullptr, nullptr, counts, nviol_accum, nullptr, upd_counts_E,
                  upd_counts_R, ws, ws_bytes, as_stream(stream), Ehat, Rhat);
}

}  // extern "C"
