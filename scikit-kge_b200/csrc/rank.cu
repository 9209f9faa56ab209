// Filtered ranking: query construction and the fp64 settlement of band candidates and
// filter entries (the coarse passes live in rank_sweep.cu, rank_umma.cu, rank_refine.cu, rank_single.cu).
//   FilteredRankingEval.positions : skge/base.py:913-1031
//   TransEEval                    : skge/run_transe.py:13-29 (always L1)
//   HolEEval                      : skge/run_hole.py:10-19
//   RESCAL (no reference evaluator): from skge/rescal.py:31-35
#include "common.cuh"

namespace skge {

// Exact (fp64) score of one entity row against one query vector; all 32 lanes
// call it and all get the result.  Used for target scores, band candidates and
// filter entries alike, so equal inputs give bit-equal scores on every GPU.
// Summation order: for d % 128 == 0 lane l owns the elements 128 j + 4 l + v (one 128-bit load of the
// entity row and two of the query row per j), v inner, j outer; otherwise the elements l + 32 j.  The
// settlement kernel below forms its sums in exactly this order.
__device__ __forceinline__ double score64_warp(int op, const double *__restrict__ q,
                                               const float *__restrict__ e, int d, int lane) {
  double acc = 0.0;
  if ((d & 127) == 0) {
    for (int c = 4 * lane; c < d; c += 128) {
      const float4 ev = __ldg(reinterpret_cast<const float4 *>(e + c));
      const double2 q01 = *reinterpret_cast<const double2 *>(q + c), q23 = *reinterpret_cast<const double2 *>(q + c + 2);
      if (op == SKGE_RANK_L1) {
        acc += fabs((double)ev.x - q01.x);
        acc += fabs((double)ev.y - q01.y);
        acc += fabs((double)ev.z - q23.x);
        acc += fabs((double)ev.w - q23.y);
      } else {
        acc = fma((double)ev.x, q01.x, acc);
        acc = fma((double)ev.y, q01.y, acc);
        acc = fma((double)ev.z, q23.x, acc);
        acc = fma((double)ev.w, q23.y, acc);
      }
    }
    return op == SKGE_RANK_L1 ? -warp_sum(acc) : warp_sum(acc);
  }
  if (op == SKGE_RANK_L1) {
    for (int c = lane; c < d; c += 32) acc += fabs((double)__ldg(e + c) - q[c]);
    return -warp_sum(acc);
  }
  for (int c = lane; c < d; c += 32) acc = fma((double)__ldg(e + c), q[c], acc);
  return warp_sum(acc);
}

// ---------------------------------------------------------------------------
// query vectors
// ---------------------------------------------------------------------------
// One CTA per query.  q64 is exact up to fp64 rounding:
//   TransE tail: E[s]+R[p]           head: E[o]-R[p]                (run_transe.py:15-29)
//   HolE   tail: cconv(R[p],E[s])    head: ccorr(R[p],E[o])         (run_hole.py:12-19)
//   RESCAL tail: W[p]^T E[s]         head: W[p] E[o]                (rescal.py:31-35)
__global__ void make_queries_kernel(int model, const float *__restrict__ E, const float *__restrict__ RW,
                                    const uint8_t *__restrict__ kind, const int32_t *__restrict__ given,
                                    const int32_t *__restrict__ rel, const int32_t *__restrict__ target,
                                    int64_t Q, int d, float enorm_max, float coarse_rel,
                                    double *__restrict__ q64, float *__restrict__ q32,
                                    double *__restrict__ tscore, float *__restrict__ eps,
                                    float *__restrict__ qnorm) {
  extern __shared__ double smd[];
  double *gv = smd;       // given entity row
  double *rv = smd + d;   // relation row (TransE / HolE)
  __shared__ double red[34];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int op = model == SKGE_MODEL_TRANSE ? SKGE_RANK_L1 : SKGE_RANK_DOT;
  for (int64_t qi = blockIdx.x; qi < Q; qi += gridDim.x) {
    __syncthreads();
    const int head = kind[qi];
    const float *eg = E + (int64_t)given[qi] * d;
    for (int i = threadIdx.x; i < d; i += blockDim.x) gv[i] = (double)__ldg(eg + i);
    if (model != SKGE_MODEL_RESCAL) {
      const float *rp = RW + (int64_t)rel[qi] * d;
      for (int i = threadIdx.x; i < d; i += blockDim.x) rv[i] = (double)__ldg(rp + i);
    }
    __syncthreads();
    double *qo = q64 + qi * d;
    double l1 = 0.0, l2 = 0.0;
    if (model == SKGE_MODEL_TRANSE) {
      for (int k = threadIdx.x; k < d; k += blockDim.x) {
        double v = head ? gv[k] - rv[k] : gv[k] + rv[k];
        qo[k] = v;
        q32[qi * d + k] = (float)v;
        l1 += fabs(v);
        l2 += v * v;
      }
    } else if (model == SKGE_MODEL_HOLE) {
      for (int k = threadIdx.x; k < d; k += blockDim.x) {
        double acc = 0.0;
        if (head) {  // ccorr(r, o)_k = sum_i r_i o_{(i+k) mod d}
          int j = k;
          for (int i = 0; i < d; ++i) {
            acc = fma(rv[i], gv[j], acc);
            if (++j == d) j = 0;
          }
        } else {     // cconv(r, s)_k = sum_i r_i s_{(k-i) mod d}
          int j = k;
          for (int i = 0; i < d; ++i) {
            acc = fma(rv[i], gv[j], acc);
            if (--j < 0) j = d - 1;
          }
        }
        qo[k] = acc;
        q32[qi * d + k] = (float)acc;
        l1 += fabs(acc);
        l2 += acc * acc;
      }
    } else {
      const float *w = RW + (int64_t)rel[qi] * d * d;
      if (head) {  // q_i = sum_j W_ij o_j : warp per row, lanes over j
        for (int r = wid; r < d; r += nw) {
          double acc = 0.0;
          for (int j = lane; j < d; j += 32) acc = fma((double)__ldg(w + (int64_t)r * d + j), gv[j], acc);
          acc = warp_sum(acc);
          if (lane == 0) {
            qo[r] = acc;
            q32[qi * d + r] = (float)acc;
            l1 += fabs(acc);
            l2 += acc * acc;
          }
        }
      } else {     // q_j = sum_i W_ij s_i : thread per column j
        for (int j = threadIdx.x; j < d; j += blockDim.x) {
          double acc = 0.0;
          for (int i = 0; i < d; ++i) acc = fma((double)__ldg(w + (int64_t)i * d + j), gv[i], acc);
          qo[j] = acc;
          q32[qi * d + j] = (float)acc;
          l1 += fabs(acc);
          l2 += acc * acc;
        }
      }
    }
    l1 = warp_sum(l1);
    l2 = warp_sum(l2);
    if (lane == 0) { red[wid] = l1; red[17 + wid] = l2; }
    __syncthreads();  // also publishes qo[] to warp 0
    if (wid == 0) {
      double a = lane < nw ? red[lane] : 0.0, b = lane < nw ? red[17 + lane] : 0.0;
      a = warp_sum(a);
      b = warp_sum(b);
      double t = score64_warp(op, qo, E + (int64_t)target[qi] * d, d, lane);
      if (lane == 0) {
        tscore[qi] = t;
        float n2 = (float)sqrt(b);
        qnorm[qi] = n2;
        eps[qi] = op == SKGE_RANK_L1 ? coarse_rel * (float)(fabs(t) + a / (double)(d + 2))
                                     : coarse_rel * n2 * enorm_max;
      }
    }
  }
}

// HolE query vectors for d % 8 == 0, d <= 256: one WARP per query, lane l owns outputs
// k = 8l .. 8l+7 (lanes beyond d/8 idle).  The correlation slides a 15-wide register window
// over a doubled copy of the given entity row in shared memory: per block of 8 inner steps a
// lane issues 8 + 8 shared loads for 64 DFMAs, so the loop runs at the FP64 pipe's pace.
// Lane l reads the window at 8 l + c: with a plain layout that is a 64-byte lane stride and a
// 16-way bank conflict, so the doubled row is stored with one pad slot per 8 (index i lives
// at i + i / 8: lane stride 72 bytes, conflict-free for 64-bit loads).
//   tail: q_k = cconv(r, s)_k = sum_i r_i s_{(k-i) mod d}     head: q_k = ccorr(r, o)_k = sum_i r_i o_{(i+k) mod d}
__global__ void __launch_bounds__(256) make_queries_hole_kernel(const float *__restrict__ E,
                                                                const float *__restrict__ R,
                                                                const uint8_t *__restrict__ kind,
                                                                const int32_t *__restrict__ given,
                                                                const int32_t *__restrict__ rel,
                                                                const int32_t *__restrict__ target, int64_t Q,
                                                                int d, float enorm_max, float coarse_rel,
                                                                double *__restrict__ q64, float *__restrict__ q32,
                                                                double *__restrict__ tscore, float *__restrict__ eps,
                                                                float *__restrict__ qnorm) {
  extern __shared__ double smd[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int g2len = 2 * d + d / 4 + 8;                // padded doubled row
  double *g2 = smd + (size_t)wid * (g2len + d);      // doubled given row, index i at HP(i)
  double *rv = g2 + g2len;                           // [d] relation row
#define HP(i) ((i) + ((i) >> 3))
  const int k0 = lane * 8;
  for (int64_t qi = (int64_t)blockIdx.x * nw + wid; qi < Q; qi += (int64_t)gridDim.x * nw) {
    const int head = kind[qi];
    const float *eg = E + (int64_t)given[qi] * d, *rp = R + (int64_t)rel[qi] * d;
    __syncwarp();
    for (int i = lane; i < d; i += 32) {
      double v = (double)__ldg(eg + i);
      g2[HP(i)] = v;
      g2[HP(i + d)] = v;
      rv[i] = (double)__ldg(rp + i);
    }
    __syncwarp();
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double l1 = 0.0, l2 = 0.0;
    if (k0 < d) {
      double w[15];
      if (head) {
        // acc[m] += r[i] * g2[i + k0 + m]; block b covers i = 8b..8b+7: window v[t] = g2[8b + k0 + t], t < 15
#pragma unroll
        for (int t = 0; t < 7; ++t) w[t] = g2[HP(k0 + t)];
        for (int b = 0; b < d / 8; ++b) {
#pragma unroll
          for (int t = 7; t < 15; ++t) w[t] = g2[HP(8 * b + k0 + t)];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const double r = rv[8 * b + u];
#pragma unroll
            for (int m = 0; m < 8; ++m) acc[m] = fma(r, w[u + m], acc[m]);
          }
#pragma unroll
          for (int t = 0; t < 7; ++t) w[t] = w[t + 8];
        }
      } else {
        // acc[m] += r[i] * g2[d + k0 + m - i]; window v[t] = g2[d + k0 - 8b - 7 + t], t < 15
#pragma unroll
        for (int t = 8; t < 15; ++t) w[t] = g2[HP(d + k0 - 7 + t)];
        for (int b = 0; b < d / 8; ++b) {
#pragma unroll
          for (int t = 0; t < 8; ++t) w[t] = g2[HP(d + k0 - 8 * b - 7 + t)];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const double r = rv[8 * b + u];
#pragma unroll
            for (int m = 0; m < 8; ++m) acc[m] = fma(r, w[m - u + 7], acc[m]);
          }
#pragma unroll
          for (int t = 14; t >= 8; --t) w[t] = w[t - 8];
        }
      }
      double *qo = q64 + qi * d + k0;
      float *qf = q32 + qi * d + k0;
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        qo[m] = acc[m];
        qf[m] = (float)acc[m];
        l1 += fabs(acc[m]);
        l2 += acc[m] * acc[m];
      }
    }
    l1 = warp_sum(l1);
    l2 = warp_sum(l2);
    __syncwarp();  // q64 row written by this warp is read back below
    double t = score64_warp(SKGE_RANK_DOT, q64 + qi * d, E + (int64_t)target[qi] * d, d, lane);
    if (lane == 0) {
      tscore[qi] = t;
      float n2 = (float)sqrt(l2);
      qnorm[qi] = n2;
      eps[qi] = coarse_rel * n2 * enorm_max;
    }
  }
#undef HP
}

// (The fp32 coarse sweep lives in rank_sweep.cu.)

// ---------------------------------------------------------------------------
// fp64 settlement
// ---------------------------------------------------------------------------
// Four (query, entity) pairs by one warp.  NC > 0: d == 32 * NC, everything unrolled -- the 4 * NC
// entity loads are issued before the first use, and when the four pairs share the query (the
// lists are sorted by query) its row is read once.  NC == 0: any d.  Each pair's sum is formed
// in score64_warp's order (c = lane, lane + 32, ...; then the warp tree), so the result is
// bit-identical to the routine that produced the target score.
template <int NC>
__device__ __forceinline__ void settle4(int op, const double *__restrict__ q64, const float *__restrict__ E, int d,
                                        const int (&qs)[4], const int (&es)[4], bool same, int lane,
                                        double (&acc)[4]) {
  const double *qp[4];
  const float *ep[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    qp[k] = q64 + (int64_t)qs[k] * d;
    ep[k] = E + (int64_t)es[k] * d;
    acc[k] = 0.0;
  }
  auto add4 = [&](double &a, const float4 &ev, const double2 &q01, const double2 &q23) {
    if (op == SKGE_RANK_L1) {
      a += fabs((double)ev.x - q01.x);
      a += fabs((double)ev.y - q01.y);
      a += fabs((double)ev.z - q23.x);
      a += fabs((double)ev.w - q23.y);
    } else {
      a = fma((double)ev.x, q01.x, a);
      a = fma((double)ev.y, q01.y, a);
      a = fma((double)ev.z, q23.x, a);
      a = fma((double)ev.w, q23.y, a);
    }
  };
  if (NC > 0 && NC % 4 == 0) {
    // d = 128 or 256: score64_warp's 128-bit order, everything unrolled, all entity loads up front
    constexpr int NJ = NC >= 4 ? NC / 4 : 1;
    float4 ev[4][NJ];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int j = 0; j < NJ; ++j) ev[k][j] = __ldg(reinterpret_cast<const float4 *>(ep[k] + 4 * lane + 128 * j));
    if (same) {
      double2 q01[NJ], q23[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        q01[j] = *reinterpret_cast<const double2 *>(qp[0] + 4 * lane + 128 * j);
        q23[j] = *reinterpret_cast<const double2 *>(qp[0] + 4 * lane + 128 * j + 2);
      }
#pragma unroll
      for (int j = 0; j < NJ; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) add4(acc[k], ev[k][j], q01[j], q23[j]);
    } else {
#pragma unroll
      for (int j = 0; j < NJ; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const double2 q01 = *reinterpret_cast<const double2 *>(qp[k] + 4 * lane + 128 * j);
          const double2 q23 = *reinterpret_cast<const double2 *>(qp[k] + 4 * lane + 128 * j + 2);
          add4(acc[k], ev[k][j], q01, q23);
        }
    }
  } else if (NC > 0) {
    constexpr int NCC = NC > 0 ? NC : 1;
    float ev[4][NCC];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int j = 0; j < NCC; ++j) ev[k][j] = __ldg(ep[k] + lane + 32 * j);
    if (same) {
      double qv[NCC];
#pragma unroll
      for (int j = 0; j < NCC; ++j) qv[j] = qp[0][lane + 32 * j];
#pragma unroll
      for (int j = 0; j < NCC; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          acc[k] = op == SKGE_RANK_L1 ? acc[k] + fabs((double)ev[k][j] - qv[j]) : fma((double)ev[k][j], qv[j], acc[k]);
    } else {
#pragma unroll
      for (int j = 0; j < NCC; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const double qv = qp[k][lane + 32 * j];
          acc[k] = op == SKGE_RANK_L1 ? acc[k] + fabs((double)ev[k][j] - qv) : fma((double)ev[k][j], qv, acc[k]);
        }
    }
  } else if ((d & 127) == 0) {
    for (int c = 4 * lane; c < d; c += 128)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 ev = __ldg(reinterpret_cast<const float4 *>(ep[k] + c));
        const double2 q01 = *reinterpret_cast<const double2 *>(qp[k] + c);
        const double2 q23 = *reinterpret_cast<const double2 *>(qp[k] + c + 2);
        add4(acc[k], ev, q01, q23);
      }
  } else {
    for (int c = lane; c < d; c += 32)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const double qv = qp[k][c];
        acc[k] = op == SKGE_RANK_L1 ? acc[k] + fabs((double)__ldg(ep[k] + c) - qv) : fma((double)__ldg(ep[k] + c), qv, acc[k]);
      }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) acc[k] = op == SKGE_RANK_L1 ? -warp_sum(acc[k]) : warp_sum(acc[k]);
}

template <int NC>
__global__ void __launch_bounds__(256, 4) rank_rescore_kernel(int op, const float *__restrict__ E, int d,
                                                           const double *__restrict__ q64,
                                                           const double *__restrict__ tscore,
                                                           const int32_t *__restrict__ pair_q,
                                                           const int32_t *__restrict__ pair_e, int64_t npairs,
                                                           const unsigned long long *__restrict__ npairs_dev,
                                                           const int32_t *__restrict__ target,
                                                           int32_t *__restrict__ cnt) {
  if (npairs_dev) {
    unsigned long long n = *npairs_dev;
    if ((int64_t)n < npairs) npairs = (int64_t)n;
  }
  if (npairs <= 0) return;
  const int lane = threadIdx.x & 31;
  int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t base = warp * 4; base < npairs; base += nwarps * 4) {
    int qs[4], es[4];
    bool live[4];
    bool same = true;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t i = base + k < npairs ? base + k : npairs - 1;
      qs[k] = pair_q[i];
      es[k] = pair_e[i];
      live[k] = base + k < npairs && !(target && target[qs[k]] == es[k]);
      same = same && qs[k] == qs[0];
    }
    double acc[4];
    settle4<NC>(op, q64, E, d, qs, es, same, lane, acc);
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (live[k] && acc[k] > tscore[qs[k]]) atomicAdd(cnt + qs[k], 1);
    }
  }
}

__global__ void __launch_bounds__(256) rank_scores_one_kernel(int op, const float *__restrict__ E, int64_t N,
                                                              int d, const double *__restrict__ q64,
                                                              double *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t e = warp; e < N; e += nwarps) {
    double s = score64_warp(op, q64, E + e * d, d, lane);
    if (lane == 0) out[e] = s;
  }
}

}  // namespace skge

using namespace skge;

extern "C" {

int skge_rank_make_queries(int model, const float *E, const float *RW, const uint8_t *kind,
                           const int32_t *given, const int32_t *rel, const int32_t *target,
                           int64_t Q, int d, float enorm_max, float coarse_rel, double *q64,
                           float *q32, double *tscore, float *eps, float *qnorm,
                           skge_stream_t stream) {
  SKGE_REQUIRE(E && RW && kind && given && rel && target && q64 && q32 && tscore && eps && qnorm,
               "null argument");
  SKGE_REQUIRE(model >= SKGE_MODEL_TRANSE && model <= SKGE_MODEL_RESCAL && d > 0 && d <= 2048 && Q >= 0,
               "bad sizes");
  if (Q == 0) return 0;
  if (model == SKGE_MODEL_HOLE && d % 8 == 0 && d <= 256) {
    size_t smem = (size_t)8 * (3 * d + d / 4 + 8) * sizeof(double);
    SKGE_CUDA(cudaFuncSetAttribute(make_queries_hole_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = (Q + 7) / 8;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    make_queries_hole_kernel<<<(int)blocks, 256, smem, as_stream(stream)>>>(
        E, RW, kind, given, rel, target, Q, d, enorm_max, coarse_rel, q64, q32, tscore, eps, qnorm);
    SKGE_LAUNCH_CHECK();
    return 0;
  }
  int threads = (d + 31) / 32 * 32;
  threads = threads < 64 ? 64 : (threads > 512 ? 512 : threads);
  int64_t blocks = Q > kNumSMs * 16 ? kNumSMs * 16 : Q;
  size_t smem = 2 * (size_t)d * sizeof(double);
  make_queries_kernel<<<(int)blocks, threads, smem, as_stream(stream)>>>(
      model, E, RW, kind, given, rel, target, Q, d, enorm_max, coarse_rel, q64, q32, tscore, eps, qnorm);
  SKGE_LAUNCH_CHECK();
  return 0;
}

int skge_rank_rescore(int op, const float *Efull, int d, const double *q64, const double *tscore,
                      const int32_t *pair_q, const int32_t *pair_e, int64_t npairs,
                      const unsigned long long *npairs_dev, const int32_t *target, int32_t *cnt,
                      skge_stream_t stream) {
  SKGE_REQUIRE(Efull && q64 && tscore && pair_q && pair_e && cnt, "null argument");
  SKGE_REQUIRE((op == SKGE_RANK_L1 || op == SKGE_RANK_DOT) && d > 0 && npairs >= 0, "bad sizes");
  if (npairs == 0) return 0;
  int64_t blocks = (npairs + 31) / 32;  // 8 warps x 4 pairs
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
#define SKGE_RESCORE(NC)                                                                                   \
  rank_rescore_kernel<NC><<<(int)blocks, 256, 0, as_stream(stream)>>>(op, Efull, d, q64, tscore, pair_q, pair_e, \
                                                                     npairs, npairs_dev, target, cnt)
  switch (d % 32 == 0 && d <= 256 ? d / 32 : 0) {
    case 1: SKGE_RESCORE(1); break;
    case 2: SKGE_RESCORE(2); break;
    case 3: SKGE_RESCORE(3); break;
    case 4: SKGE_RESCORE(4); break;
    case 5: SKGE_RESCORE(5); break;
    case 6: SKGE_RESCORE(6); break;
    case 7: SKGE_RESCORE(7); break;
    case 8: SKGE_RESCORE(8); break;
    default: SKGE_RESCORE(0); break;
  }
#undef SKGE_RESCORE
  SKGE_LAUNCH_CHECK();
  return 0;
}

int skge_rank_scores_one(int op, const float *E, int64_t N, int d, const double *q64, double *out,
                         skge_stream_t stream) {
  SKGE_REQUIRE(E && q64 && out && N >= 0 && d > 0, "bad arguments");
  if (N == 0) return 0;
  int64_t blocks = (N + 7) / 8;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  rank_scores_one_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(op, E, N, d, q64, out);
  SKGE_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
