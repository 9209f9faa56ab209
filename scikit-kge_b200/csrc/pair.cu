// Pairwise-margin minibatch kernels (TransE, HolE) and their C entry points.
//   TransE._pairwise_gradients : skge/transe.py:48-165
//   HolE._pairwise_gradients   : skge/hole.py:44-100
//   PairwiseStochasticTrainer._process_batch / _batch_step : skge/base.py:1394-1427, 1306-1316
#include "common.cuh"
#include "hole_math.cuh"
#include "segment.cuh"

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
__global__ void hole_pair_kernel(const float *__restrict__ E, const float *__restrict__ R, PairIdx ix,
                                 int64_t P, int d, int af, float margin, float *__restrict__ pscores,
                                 float *__restrict__ nscores, uint8_t *__restrict__ flags,
                                 float *__restrict__ G, int32_t *__restrict__ counts,
                                 int64_t *__restrict__ nviol_accum) {
  extern __shared__ float sm[];
  // per triple: s[d], r[d], o2[2d], rrev2[2d]
  float *s_p = sm, *r_p = s_p + d, *o2_p = r_p + d, *rr_p = o2_p + 2 * d;
  float *s_n = rr_p + 2 * d, *r_n = s_n + d, *o2_n = r_n + d, *rr_n = o2_n + 2 * d;
  float *red = rr_n + 2 * d;
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
    const float *rp = R + (int64_t)ix.pp[i] * d, *rn = R + (int64_t)ix.pn[i] * d;
    smem_load(s_p, E + (int64_t)ix.sp[i] * d, d);
    smem_load(r_p, rp, d);
    smem_load_doubled(o2_p, E + (int64_t)ix.op[i] * d, d);
    smem_load_rev_doubled(rr_p, rp, d);
    smem_load(s_n, E + (int64_t)ix.sn[i] * d, d);
    smem_load(r_n, rn, d);
    smem_load_doubled(o2_n, E + (int64_t)ix.on[i] * d, d);
    smem_load_rev_doubled(rr_n, rn, d);
    __syncthreads();
    // phase 1 (d <= blockDim.x is guaranteed by the launcher)
    const int k = threadIdx.x;
    float cso_p = 0.f, cso_n = 0.f;
    if (k < d) {
      cso_p = sliding_dot(s_p, o2_p, k, d);
      cso_n = sliding_dot(s_n, o2_n, k, d);
    }
    float raw_p = block_sum(k < d ? r_p[k] * cso_p : 0.f, red);
    float raw_n = block_sum(k < d ? r_n[k] * cso_n : 0.f, red);
    float fp = act_f(af, raw_p), fn = act_f(af, raw_n);
    bool viol = fn + margin > fp;  // skge/hole.py:56
    if (threadIdx.x == 0) {
      flags[i] = viol;
      if (pscores) pscores[i] = raw_p;
      if (nscores) nscores[i] = raw_n;
      if (viol) {
        atomicAdd(counts, 1);
        if (nviol_accum) atomicAdd(reinterpret_cast<unsigned long long *>(nviol_accum), 1ull);
      }
    }
    if (!viol || k >= d) continue;
    float gp = -act_g_given_f(af, fp), gn = act_g_given_f(af, fn);  // skge/hole.py:66-67
    float *g = G + (int64_t)i * 6 * d;
    int offc = (d - k) % d;
    g[0 * d + k] = gp * sliding_dot(r_p, o2_p, k, d);
    g[1 * d + k] = gn * sliding_dot(r_n, o2_n, k, d);
    g[2 * d + k] = gp * sliding_dot(s_p, rr_p, offc, d);
    g[3 * d + k] = gn * sliding_dot(s_n, rr_n, offc, d);
    g[4 * d + k] = gp * cso_p;
    g[5 * d + k] = gn * cso_n;
  }
}

static int pair_block_threads(int d) {
  int t = (d + 31) / 32 * 32;
  return t < 64 ? 64 : t;
}

struct PairBuffers {
  uint8_t *flags;
  float *G;
};

static size_t pair_ws_bytes(int64_t P, int d, int rows, int nroles) {
  if (P < 1) P = 1;
  return align_up((size_t)P) + align_up((size_t)P * rows * d * sizeof(float)) +
         seg_workspace_bytes((int64_t)nroles * P, d) + 1024;
}

// model: 0 TransE, 1 HolE
static int pair_run(int model, float *E, float *R, float *p2E, float *p2R, const PairIdx &ix, int64_t P,
                    int64_t N, int64_t M, int d, int l1_or_af, float margin, float rparam, bool update,
                    int opt, float lr, int postE, int postR, float *pscores, float *nscores, float *ge,
                    int32_t *eidx, float *gr, int32_t *ridx, int32_t *counts, int64_t *nviol_accum,
                    int32_t *ent_viol, int32_t *ucE, int32_t *ucR, void *ws, size_t ws_bytes,
                    cudaStream_t st) {
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
  if (model == 0) {
    int64_t blocks = (P + 7) / 8;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    switch (pick_vec(d)) {
      case 4:
        transe_pair_kernel<4><<<(int)blocks, 256, 0, st>>>(E, R, ix, P, d, l1_or_af, margin, pscores, nscores,
                                                          flags, G, counts, nviol_accum, ent_viol);
        break;
      case 2:
        transe_pair_kernel<2><<<(int)blocks, 256, 0, st>>>(E, R, ix, P, d, l1_or_af, margin, pscores, nscores,
                                                          flags, G, counts, nviol_accum, ent_viol);
        break;
      default:
        transe_pair_kernel<1><<<(int)blocks, 256, 0, st>>>(E, R, ix, P, d, l1_or_af, margin, pscores, nscores,
                                                          flags, G, counts, nviol_accum, ent_viol);
        break;
    }
  } else {
    SKGE_REQUIRE(d <= 1024, "HolE pair kernel supports d <= 1024");
    size_t smem = (12 * (size_t)d + 40) * sizeof(float);
    SKGE_CUDA(cudaFuncSetAttribute(hole_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = P > kNumSMs * 16 ? kNumSMs * 16 : P;
    hole_pair_kernel<<<(int)blocks, pair_block_threads(d), smem, st>>>(E, R, ix, P, d, l1_or_af, margin,
                                                                      pscores, nscores, flags, G, counts,
                                                                      nviol_accum);
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
    for (int r = 0; r < 6; ++r) { rm.idx[r] = idx[r]; rm.is_rel[r] = isrel[r]; rm.grow[r] = grow[r]; rm.gsign[r] = 1.f; }
  }
  rm.nroles = 6;
  ParamDesc pd[2];
  pd[0] = ParamDesc{E, p2E, postE, 0.f, ucE, ge, eidx};
  pd[1] = ParamDesc{R, p2R, postR, rparam, ucR, gr, ridx};
  return seg_run(rm, flags, P, N, M, d, G, rows, pd, update, opt, lr, counts, ar, st);
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

}  // extern "C"
