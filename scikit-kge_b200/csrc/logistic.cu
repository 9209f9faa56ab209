// Logistic-loss minibatch kernels (HolE, RESCAL) and their C entry points.
//   HolE._gradients   : skge/hole.py:22-42
//   RESCAL._gradients : skge/rescal.py:37-76
//   StochasticTrainer._process_batch / _batch_step : skge/base.py:1293-1316
#include "common.cuh"
#include "hole_math.cuh"
#include "segment.cuh"

namespace skge {

// loss term logaddexp(0, -ys) and fs = -(y * sigmoid(-ys))  (skge/hole.py:26-28)
__device__ __forceinline__ void logistic_terms(float y, float raw, float *loss, float *fs) {
  float ys = y * raw;
  float m = -ys;
  *loss = fmaxf(m, 0.f) + log1pf(expf(-fabsf(m)));
  *fs = -(y / (1.0f + expf(ys)));
}

// One CTA per example; thread k owns component k.
//   G[i][0] = fs ccorr(R[p],E[o]) -> s   G[i][1] = fs cconv(E[s],R[p]) -> o   G[i][2] = fs ccorr(E[s],E[o]) -> p
__global__ void hole_logistic_kernel(const float *__restrict__ E, const float *__restrict__ R,
                                     const int32_t *__restrict__ s, const int32_t *__restrict__ o,
                                     const int32_t *__restrict__ p, const float *__restrict__ y,
                                     const uint8_t *__restrict__ valid, int64_t n, int d, float *__restrict__ G,
                                     double *__restrict__ loss, double *__restrict__ loss_accum,
                                     int32_t *__restrict__ counts) {
  extern __shared__ float sm[];
  float *es = sm, *rp = es + d, *o2 = rp + d, *rr = o2 + 2 * d, *red = rr + 2 * d;
  double lsum = 0.0;
  int nvalid = 0;
  for (int64_t i = blockIdx.x; i < n; i += gridDim.x) {
    if (valid && !valid[i]) continue;   // a negative the sampler could not produce (skge/sample.py:22-24)
    ++nvalid;
    __syncthreads();
    const float *rg = R + (int64_t)p[i] * d;
    smem_load(es, E + (int64_t)s[i] * d, d);
    smem_load(rp, rg, d);
    smem_load_doubled(o2, E + (int64_t)o[i] * d, d);
    smem_load_rev_doubled(rr, rg, d);
    __syncthreads();
    const int k = threadIdx.x;
    float cso = k < d ? sliding_dot(es, o2, k, d) : 0.f;
    float raw = block_sum(k < d ? rp[k] * cso : 0.f, red);
    float l, fs;
    logistic_terms(y[i], raw, &l, &fs);
    lsum += l;
    if (k < d) {
      float *g = G + (int64_t)i * 3 * d;
      g[k] = fs * sliding_dot(rp, o2, k, d);
      g[d + k] = fs * sliding_dot(es, rr, (d - k) % d, d);
      g[2 * d + k] = fs * cso;
    }
  }
  if (threadIdx.x == 0 && nvalid) {
    if (loss) atomicAdd(loss, lsum);
    if (loss_accum) atomicAdd(loss_accum, lsum);
    atomicAdd(counts, nvalid);
  }
}

// One CTA per example.  EW_j = sum_i E[s]_i W[p]_ij (thread j, coalesced over j);
// WE_i = sum_j W[p]_ij E[o]_j (warp per row i, lanes over j).
//   G[i][0] = fs * WE -> s     G[i][1] = fs * EW -> o        (skge/rescal.py:72-73)
__global__ void rescal_logistic_kernel(const float *__restrict__ E, const float *__restrict__ W,
                                       const int32_t *__restrict__ s, const int32_t *__restrict__ o,
                                       const int32_t *__restrict__ p, const float *__restrict__ y,
                                       const uint8_t *__restrict__ valid, int64_t n, int d,
                                       float *__restrict__ G, float *__restrict__ fsv, double *__restrict__ loss,
                                       double *__restrict__ loss_accum, int32_t *__restrict__ counts) {
  extern __shared__ float sm[];
  float *es = sm, *eo = es + d, *we = eo + d, *red = we + d;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  double lsum = 0.0;
  int nvalid = 0;
  for (int64_t i = blockIdx.x; i < n; i += gridDim.x) {
    if (valid && !valid[i]) continue;
    ++nvalid;
    __syncthreads();
    smem_load(es, E + (int64_t)s[i] * d, d);
    smem_load(eo, E + (int64_t)o[i] * d, d);
    __syncthreads();
    const float *w = W + (int64_t)p[i] * d * d;
    for (int r = wid; r < d; r += nw) {
      float a = 0.f;
      for (int j = lane; j < d; j += 32) a = fmaf(__ldg(w + (int64_t)r * d + j), eo[j], a);
      a = warp_sum(a);
      if (lane == 0) we[r] = a;
    }
    float ew[4] = {0.f, 0.f, 0.f, 0.f};  // d <= 4 * blockDim.x (checked by the launcher)
    for (int r = 0; r < d; ++r) {
      float sv = es[r];
      const float *wr = w + (int64_t)r * d;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        int j = threadIdx.x + t * blockDim.x;
        if (j < d) ew[t] = fmaf(sv, __ldg(wr + j), ew[t]);
      }
    }
    __syncthreads();
    float part = 0.f;
    for (int r = threadIdx.x; r < d; r += blockDim.x) part += es[r] * we[r];
    float raw = block_sum(part, red);
    float l, fs;
    logistic_terms(y[i], raw, &l, &fs);
    lsum += l;
    float *g = G + (int64_t)i * 2 * d;
    for (int r = threadIdx.x; r < d; r += blockDim.x) g[r] = fs * we[r];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      int j = threadIdx.x + t * blockDim.x;
      if (j < d) g[d + j] = fs * ew[t];
    }
    if (threadIdx.x == 0) fsv[i] = fs;
  }
  if (threadIdx.x == 0 && nvalid) {
    if (loss) atomicAdd(loss, lsum);
    if (loss_accum) atomicAdd(loss_accum, lsum);
    atomicAdd(counts, nvalid);
  }
}

// The same per-example math for minibatches already grouped by relation (the sorted example list
// the W update needs anyway).  The one-CTA-per-example kernel above re-reads the d x d relation
// matrix from L2 twice per example (80 KB at d = 100: 145 us of a 313 us config-3 minibatch); here a
// CTA takes a chunk of EB consecutive examples of the sorted list, keeps W[p] in shared memory for the
// run of examples that share it and computes both products for the run as two small register-tiled
// GEMMs.  Warp g owns EX examples, lane q owns MR rows of W . E[o] (rows q, q + 32, ...: eight lanes of a
// 128-bit phase read eight consecutive rows, whose stride is chosen so that they fall in different bank
// groups) and MR columns of E[s] . W (columns MR q ..: contiguous 128-bit reads), so one k-step of four
// moves 4 MR + 2 EX 128-bit words through shared memory for 8 MR EX FMAs.  The score is a warp sum (a
// warp holds all rows of its examples): no block-level exchange, fs goes round by shuffle.
// The reference memoises per-relation products for the same reason (skge/rescal.py:43-57).
template <int MR>
struct RescalGrouped {
  static constexpr int EX = 16 / MR;        // examples per warp: 4 (d <= 128) or 2 (d <= 256)
  static constexpr int EB = 8 * EX;         // examples per chunk (8 warps)
  static constexpr int RUN_SPLIT = 4;       // CTAs per chunk: run r of a chunk goes to CTA r % RUN_SPLIT
  static constexpr int DMAX = 32 * MR;
  // row stride of W in shared memory: a multiple of 4 floats whose quarter is odd (see above)
  __host__ __device__ static int w_stride(int DS) { return ((DS >> 2) & 1) ? DS : DS + 4; }
  __host__ __device__ static size_t smem_bytes(int d) {
    const size_t DS = (size_t)((d + 3) & ~3);
    return (DS * (DS + 4) + 2 * EB * DS + 8) * sizeof(float) + 2 * EB * sizeof(int32_t) + 8 * sizeof(double);
  }
};

// asynchronous global -> shared copies: a thread puts all of its elements of a gather in flight at once
// (no registers); the gathers here are latency-, not bandwidth-bound
__device__ __forceinline__ void cp_async_f32(float *smem_dst, const float *gsrc) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(a), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_f32x4(float *smem_dst, const float *gsrc) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// one row of d floats -> shared memory (zero padded to DS); vec16: d % 4 == 0 and 16-byte aligned source
__device__ __forceinline__ void row_to_smem_async(float *dst, const float *src, int d, int DS, int lane, bool vec16) {
  if (vec16) {
    for (int q = lane; q < (d >> 2); q += 32) cp_async_f32x4(dst + 4 * q, src + 4 * q);
  } else {
    for (int j = lane; j < DS; j += 32) {
      if (j < d) cp_async_f32(dst + j, src + j);
      else dst[j] = 0.f;
    }
  }
}

template <int MR, bool MASKED>
__device__ __forceinline__ void rescal_run_products(const float *__restrict__ Ws, const float *__restrict__ es,
                                                    const float *__restrict__ eo, int DS, int WS, int q, int e0,
                                                    unsigned emask, float (&we)[16 / MR][MR], float (&ew)[16 / MR][MR]) {
  constexpr int EX = 16 / MR, CH = MR / 4;
  bool rok[MR], cok[CH];
#pragma unroll
  for (int m = 0; m < MR; ++m) rok[m] = q + 32 * m < DS;
#pragma unroll
  for (int h = 0; h < CH; ++h) cok[h] = MR * q + 4 * h < DS;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k = 0; k < DS; k += 4) {
    float4 wr[MR], wc[4][CH];
#pragma unroll
    for (int m = 0; m < MR; ++m)      // W[q + 32 m][k .. k + 3]
      wr[m] = rok[m] ? *reinterpret_cast<const float4 *>(Ws + (q + 32 * m) * WS + k) : zero4;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int h = 0; h < CH; ++h)    // W[k + i][MR q + 4 h .. + 3]
        wc[i][h] = cok[h] ? *reinterpret_cast<const float4 *>(Ws + (k + i) * WS + MR * q + 4 * h) : zero4;
#pragma unroll
    for (int e = 0; e < EX; ++e) {
      if (MASKED && !((emask >> e) & 1u)) continue;
      const float4 vo = *reinterpret_cast<const float4 *>(eo + (e0 + e) * DS + k);
      const float4 vs = *reinterpret_cast<const float4 *>(es + (e0 + e) * DS + k);
#pragma unroll
      for (int m = 0; m < MR; ++m)
        we[e][m] = fmaf(wr[m].w, vo.w, fmaf(wr[m].z, vo.z, fmaf(wr[m].y, vo.y, fmaf(wr[m].x, vo.x, we[e][m]))));
#pragma unroll
      for (int h = 0; h < CH; ++h) {
        ew[e][4 * h + 0] = fmaf(wc[3][h].x, vs.w, fmaf(wc[2][h].x, vs.z, fmaf(wc[1][h].x, vs.y, fmaf(wc[0][h].x, vs.x, ew[e][4 * h + 0]))));
        ew[e][4 * h + 1] = fmaf(wc[3][h].y, vs.w, fmaf(wc[2][h].y, vs.z, fmaf(wc[1][h].y, vs.y, fmaf(wc[0][h].y, vs.x, ew[e][4 * h + 1]))));
        ew[e][4 * h + 2] = fmaf(wc[3][h].z, vs.w, fmaf(wc[2][h].z, vs.z, fmaf(wc[1][h].z, vs.y, fmaf(wc[0][h].z, vs.x, ew[e][4 * h + 2]))));
        ew[e][4 * h + 3] = fmaf(wc[3][h].w, vs.w, fmaf(wc[2][h].w, vs.z, fmaf(wc[1][h].w, vs.y, fmaf(wc[0][h].w, vs.x, ew[e][4 * h + 3]))));
      }
    }
  }
}

template <int MR>
__global__ void __launch_bounds__(256) rescal_logistic_grouped_kernel(
    const float *__restrict__ E, const float *__restrict__ W, const int32_t *__restrict__ s,
    const int32_t *__restrict__ o, const int32_t *__restrict__ p, const float *__restrict__ y,
    const int32_t *__restrict__ sorted_vals, const int32_t *__restrict__ meta, int d, int vec16,
    float *__restrict__ G, float *__restrict__ fsv, double *__restrict__ loss, double *__restrict__ loss_accum,
    int32_t *__restrict__ counts) {
  using RG = RescalGrouped<MR>;
  constexpr int EX = RG::EX, EB = RG::EB, CH = MR / 4;
  extern __shared__ __align__(16) float smg[];
  const int DS = (d + 3) & ~3, WS = RG::w_stride(DS);
  float *Ws = smg;                     // [DS][WS], zero beyond d
  float *es = Ws + DS * WS;            // [EB][DS]
  float *eo = es + EB * DS;            // [EB][DS]
  double *lred = reinterpret_cast<double *>(eo + EB * DS);            // [8] per-warp loss sums (8-byte aligned)
  int32_t *exs = reinterpret_cast<int32_t *>(lred + 8);               // [EB] example ids of the chunk
  int32_t *rels = exs + EB;                                            // [EB] their relations
  const int total = meta[3];           // valid examples in the sorted list
  const int t = threadIdx.x, q = t & 31, warp = t >> 5, e0 = warp * EX;
  double lsum = 0.0;                   // lane 0 of each warp: the losses of the warp's examples
  int nval = 0;
  for (int chunk = blockIdx.x; chunk * EB < total; chunk += gridDim.x) {
    const int cbeg = chunk * EB, cn = min(EB, total - cbeg);
    __syncthreads();
    if (t < EB) {
      const int ex = t < cn ? (sorted_vals[cbeg + t] >> 4) : -1;
      exs[t] = ex;
      rels[t] = ex >= 0 ? p[ex] : -1;
    }
    __syncthreads();
    if (blockIdx.y == 0) nval += cn;
    int nruns = 1;
    for (int e = 1; e < cn; ++e) nruns += rels[e] != rels[e - 1];
    if ((int)blockIdx.y >= nruns) continue;   // this CTA owns runs y, y + RUN_SPLIT, ...: none here
    for (int row = warp; row < 2 * EB; row += 8) {
      const int e = row >> 1;
      float *dst = ((row & 1) ? eo : es) + e * DS;
      if (e < cn) {
        const int ex = exs[e];
        row_to_smem_async(dst, E + (int64_t)((row & 1) ? o[ex] : s[ex]) * d, d, DS, q, vec16 != 0);
      } else {
        for (int j = q; j < DS; j += 32) dst[j] = 0.f;
      }
    }
    int rb = 0, ri = 0;
    while (rb < cn) {   // runs of one relation inside the chunk (usually the whole chunk)
      const int rel = rels[rb];
      int re = rb + 1;
      while (re < cn && rels[re] == rel) ++re;
      if (ri++ % (int)gridDim.y != (int)blockIdx.y) { rb = re; continue; }
      __syncthreads();   // the previous run is done with Ws
      const float *w = W + (int64_t)rel * d * d;
      if (vec16) {
        for (int r = warp; r < d; r += 8)
          for (int j4 = q; j4 < (d >> 2); j4 += 32) cp_async_f32x4(Ws + r * WS + 4 * j4, w + (int64_t)r * d + 4 * j4);
      } else {
        for (int r = warp; r < DS; r += 8)
          for (int j = q; j < DS; j += 32) {
            if (r < d && j < d) cp_async_f32(Ws + r * WS + j, w + (int64_t)r * d + j);
            else Ws[r * WS + j] = 0.f;
          }
      }
      cp_async_wait_all();   // also the chunk's E rows, issued above
      __syncthreads();
      // examples of this warp that belong to the run
      const int lo = max(rb, e0) - e0, hi = min(re, e0 + EX) - e0;
      const unsigned emask = hi > lo ? ((1u << hi) - 1u) & ~((1u << lo) - 1u) : 0u;   // warp-uniform
      if (emask) {
        float we[EX][MR], ew[EX][MR];
#pragma unroll
        for (int e = 0; e < EX; ++e)
#pragma unroll
          for (int m = 0; m < MR; ++m) we[e][m] = ew[e][m] = 0.f;
        if (emask == (1u << EX) - 1u) rescal_run_products<MR, false>(Ws, es, eo, DS, WS, q, e0, emask, we, ew);
        else rescal_run_products<MR, true>(Ws, es, eo, DS, WS, q, e0, emask, we, ew);
#pragma unroll
        for (int e = 0; e < EX; ++e) {
          if (!((emask >> e) & 1u)) continue;
          // raw score = E[s] . (W E[o]): the warp holds every row of the example
          float pr = 0.f;
#pragma unroll
          for (int m = 0; m < MR; ++m)
            if (q + 32 * m < DS) pr = fmaf(es[(e0 + e) * DS + q + 32 * m], we[e][m], pr);
          const float raw = warp_sum(pr);
          const int ex = exs[e0 + e];
          float l, fs;
          logistic_terms(y[ex], raw, &l, &fs);
          if (q == 0) {
            fsv[ex] = fs;
            lsum += (double)l;
          }
          float *gr = G + (int64_t)ex * 2 * d;
#pragma unroll
          for (int m = 0; m < MR; ++m)
            if (q + 32 * m < d) gr[q + 32 * m] = fs * we[e][m];            // -> s   (skge/rescal.py:72)
#pragma unroll
          for (int h = 0; h < CH; ++h) {                                   // -> o   (skge/rescal.py:73)
            const int c0 = MR * q + 4 * h;
            if (vec16) {
              if (c0 < d)
                *reinterpret_cast<float4 *>(gr + d + c0) =
                    make_float4(fs * ew[e][4 * h], fs * ew[e][4 * h + 1], fs * ew[e][4 * h + 2], fs * ew[e][4 * h + 3]);
            } else {
#pragma unroll
              for (int v = 0; v < 4; ++v)
                if (c0 + v < d) gr[d + c0 + v] = fs * ew[e][4 * h + v];
            }
          }
        }
      }
      rb = re;
    }
  }
  // losses: per-warp sums combined in warp order
  __syncthreads();
  if (q == 0) lred[warp] = lsum;
  __syncthreads();
  if (t == 0) {
    double tot = 0.0;
    for (int w8 = 0; w8 < 8; ++w8) tot += lred[w8];
    if (loss && tot != 0.0) atomicAdd(loss, tot);
    if (loss_accum && tot != 0.0) atomicAdd(loss_accum, tot);
    if (nval) atomicAdd(counts, nval);
  }
}

// gw[u] = mean_{i in relation u} fs_i E[s_i] E[o_i]^T + rparam W[p_u]   (skge/rescal.py:61-70)
// A frequent relation owns a large share of the minibatch, so its examples are cut into
// kGwSlices slices reduced by different CTAs into partial sums (pass 1, grid = (tiles_b,
// tiles_a, segments * slices); 32x32 output tile per CTA, 256 threads x 4 outputs); pass 2
// adds the slices in a fixed order, takes the mean and adds the regulariser.
static constexpr int kGwSlices = 8;

__global__ void __launch_bounds__(256) rescal_gw_partial_kernel(const float *__restrict__ E,
                                                                const int32_t *__restrict__ s,
                                                                const int32_t *__restrict__ o,
                                                                const float *__restrict__ fsv, SegLists sl, int d,
                                                                float *__restrict__ part, int maxseg) {
  __shared__ float ts[32][33], to[32][33], tf[32];
  __shared__ int32_t srow[32], orow[32];
  const int nseg = sl.meta[0];
  const int a0 = blockIdx.y * 32, b0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // ty in 0..7, rows ty, ty+8, ty+16, ty+24
  for (int item = blockIdx.z; item < nseg * kGwSlices; item += gridDim.z) {
    const int seg = item / kGwSlices, slice = item - seg * kGwSlices;
    const int sbeg = sl.seg_start[seg], len = sl.seg_start[seg + 1] - sbeg;
    const int beg = sbeg + (int)((int64_t)len * slice / kGwSlices);
    const int end = sbeg + (int)((int64_t)len * (slice + 1) / kGwSlices);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j0 = beg; j0 < end; j0 += 32) {
      int cnt = min(32, end - j0);
      __syncthreads();
      // the batch's example ids and row numbers first, so that the tile loads below are independent
      // (payload -> id -> row was a chain of three dependent loads per element)
      if (threadIdx.x < 32) {
        const int ex = threadIdx.x < cnt ? (sl.vals[j0 + threadIdx.x] >> 4) : -1;
        srow[threadIdx.x] = ex >= 0 ? s[ex] : 0;
        orow[threadIdx.x] = ex >= 0 ? o[ex] : 0;
        tf[threadIdx.x] = ex >= 0 ? fsv[ex] : 0.f;
      }
      __syncthreads();
      float vs[4], vo[4];
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int e = ty + 8 * it;
        vs[it] = (e < cnt && a0 + tx < d) ? __ldg(E + (int64_t)srow[e] * d + a0 + tx) : 0.f;
        vo[it] = (e < cnt && b0 + tx < d) ? __ldg(E + (int64_t)orow[e] * d + b0 + tx) : 0.f;
      }
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        ts[ty + 8 * it][tx] = vs[it];
        to[ty + 8 * it][tx] = vo[it];
      }
      __syncthreads();
      for (int e = 0; e < cnt; ++e) {
        float fo = tf[e] * to[e][tx];
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q] = fmaf(ts[e][ty + 8 * q], fo, acc[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int a = a0 + ty + 8 * q, b = b0 + tx;
      if (a < d && b < d) part[(((int64_t)slice * maxseg + seg) * d + a) * d + b] = acc[q];
    }
  }
}

__global__ void __launch_bounds__(256) rescal_gw_finish_kernel(const float *__restrict__ W, SegLists sl, int d,
                                                               float rparam, const float *__restrict__ part,
                                                               int maxseg, float *__restrict__ gw,
                                                               int32_t *__restrict__ pidx,
                                                               int32_t *__restrict__ counts) {
  const int nseg = sl.meta[0];
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && counts) counts[2] = nseg;
  const int64_t dd = (int64_t)d * d;
  for (int seg = blockIdx.y; seg < nseg; seg += gridDim.y) {
    const int rel = sl.seg_key[seg];
    const float inv = 1.0f / (float)(sl.seg_start[seg + 1] - sl.seg_start[seg]);
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < dd; c += (int64_t)gridDim.x * blockDim.x) {
      float v = 0.f;
#pragma unroll
      for (int k = 0; k < kGwSlices; ++k) v += part[((int64_t)k * maxseg + seg) * dd + c];
      v *= inv;
      if (rparam != 0.f) v += rparam * __ldg(W + (int64_t)rel * dd + c);
      gw[(int64_t)seg * dd + c] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) pidx[seg] = rel;
  }
}

static int block_threads(int d) {
  int t = (d + 31) / 32 * 32;
  return t < 64 ? 64 : (t > 1024 ? 1024 : t);
}

static size_t logistic_ws_bytes(int model, int64_t n, int d, int64_t M) {
  if (n < 1) n = 1;
  int rows = model == SKGE_MODEL_RESCAL ? 2 : 3;
  size_t b = align_up((size_t)n * rows * d * sizeof(float));
  b += seg_workspace_bytes((int64_t)3 * n, d);
  if (model == SKGE_MODEL_RESCAL) {
    int64_t uw = n < M ? n : M;
    b += align_up((size_t)n * sizeof(float));                     // fs
    b += seg_workspace_bytes(n, 0);                               // relation grouping
    b += align_up((size_t)uw * d * d * sizeof(float));            // gw (step mode)
    b += align_up((size_t)uw * sizeof(int32_t));                  // pidx (step mode)
    b += align_up((size_t)kGwSlices * uw * d * d * sizeof(float));  // per-slice partial sums of gw
  }
  return b + 1024;
}

static int hole_logistic_run(float *E, float *R, float *p2E, float *p2R, const int32_t *s, const int32_t *o,
                             const int32_t *p, const float *y, const uint8_t *valid, int64_t n, int64_t N, int64_t M, int d,
                             float rparam, bool update, int opt, float lr, int postE, int postR, float *ge,
                             int32_t *eidx, float *gr, int32_t *ridx, int32_t *counts, double *loss,
                             double *loss_accum, int32_t *ucE, int32_t *ucR, void *ws, size_t ws_bytes,
                             cudaStream_t st) {
  SKGE_REQUIRE(E && R && s && o && p && y && counts && ws, "null argument");
  SKGE_REQUIRE(n > 0 && d > 0 && d <= 1024 && N > 0 && M > 0, "bad sizes");
  if (update) SKGE_REQUIRE(opt == SKGE_OPT_SGD || (p2E && p2R), "AdaGrad needs p2E/p2R");
  else SKGE_REQUIRE(ge && eidx && gr && ridx, "null output");
  Arena ar(ws, ws_bytes);
  float *G = ar.take<float>((size_t)n * 3 * d);
  if (!ar.ok()) {
    set_error("workspace too small");
    return SKGE_EWORKSPACE;
  }
  SKGE_CUDA(cudaMemsetAsync(counts, 0, 4 * sizeof(int32_t), st));
  if (loss) SKGE_CUDA(cudaMemsetAsync(loss, 0, sizeof(double), st));
  size_t smem = (6 * (size_t)d + 40) * sizeof(float);
  SKGE_CUDA(cudaFuncSetAttribute(hole_logistic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t blocks = n > kNumSMs * 16 ? kNumSMs * 16 : n;
  hole_logistic_kernel<<<(int)blocks, block_threads(d), smem, st>>>(E, R, s, o, p, y, valid, n, d, G, loss,
                                                                    loss_accum, counts);
  SKGE_LAUNCH_CHECK();
  RoleMap rm;
  const int32_t *idx[3] = {s, o, p};  // entity keys ss+os, relation keys ps: hole.py:31-39
  for (int r = 0; r < 3; ++r) { rm.idx[r] = idx[r]; rm.is_rel[r] = r == 2; rm.grow[r] = r; rm.gsign[r] = 1.f; }
  rm.nroles = 3;
  ParamDesc pd[2];
  pd[0] = ParamDesc{E, p2E, postE, rparam, ucE, ge, eidx};
  pd[1] = ParamDesc{R, p2R, postR, rparam, ucR, gr, ridx};
  return seg_run(rm, valid, n, N, M, d, G, 3, pd, update, opt, lr, counts, ar, st);
}

// sparse update whose row count lives on the device: run over the maximum and
// let rows >= *U_dev exit (idx rows beyond U are never touched).
__global__ void __launch_bounds__(256) rescal_w_update_kernel(float *W, float *p2, const float *__restrict__ gw,
                                                              const int32_t *__restrict__ pidx,
                                                              const int32_t *__restrict__ counts, int64_t rowlen,
                                                              int opt, float lr, int32_t *upd_counts) {
  int U = counts[2];
  for (int u = blockIdx.y; u < U; u += gridDim.y) {
    int64_t row = pidx[u];
    float *x = W + row * rowlen;
    float *a2 = p2 ? p2 + row * rowlen : nullptr;
    const float *g = gw + (int64_t)u * rowlen;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < rowlen; c += (int64_t)gridDim.x * blockDim.x) {
      float gg = __ldg(g + c), xv = x[c];
      if (opt == SKGE_OPT_ADAGRAD) {
        float a = a2[c] + gg * gg;
        a2[c] = a;
        xv -= lr * gg * adagrad_rscale(a);
      } else {
        xv -= lr * gg;
      }
      x[c] = xv;
    }
    if (upd_counts && blockIdx.x == 0 && threadIdx.x == 0) upd_counts[row] += 1;
  }
}

static int rescal_logistic_run(float *E, float *W, float *p2E, float *p2W, const int32_t *s, const int32_t *o,
                               const int32_t *p, const float *y, const uint8_t *valid, int64_t n, int64_t N, int64_t M, int d,
                               float rparam, bool update, int opt, float lr, int postE, int postW, float *ge,
                               int32_t *eidx, float *gw, int32_t *pidx, int32_t *counts, double *loss,
                               double *loss_accum, int32_t *ucE, int32_t *ucW, void *ws, size_t ws_bytes,
                               cudaStream_t st) {
  SKGE_REQUIRE(E && W && s && o && p && y && counts && ws, "null argument");
  SKGE_REQUIRE(n > 0 && d > 0 && d <= 1024 && N > 0 && M > 0, "bad sizes");
  SKGE_REQUIRE(postW == SKGE_POST_NONE, "post-hooks on W are not supported");
  if (update) SKGE_REQUIRE(opt == SKGE_OPT_SGD || (p2E && p2W), "AdaGrad needs p2E/p2W");
  else SKGE_REQUIRE(ge && eidx && gw && pidx, "null output");
  Arena ar(ws, ws_bytes);
  float *G = ar.take<float>((size_t)n * 2 * d);
  float *fsv = ar.take<float>(n);
  int64_t uw = n < M ? n : M;
  if (update) {
    gw = ar.take<float>((size_t)uw * d * d);
    pidx = ar.take<int32_t>(uw);
  }
  float *gw_part = ar.take<float>((size_t)kGwSlices * uw * d * d);
  if (!ar.ok()) {
    set_error("workspace too small");
    return SKGE_EWORKSPACE;
  }
  SKGE_CUDA(cudaMemsetAsync(counts, 0, 4 * sizeof(int32_t), st));
  if (loss) SKGE_CUDA(cudaMemsetAsync(loss, 0, sizeof(double), st));
  // group examples by relation first: the per-relation outer-product mean needs the lists, and the
  // grouped logistic kernel keeps W[p] in shared memory over a run of them
  RoleMap rmw;
  rmw.idx[0] = p; rmw.is_rel[0] = 0; rmw.grow[0] = 0; rmw.gsign[0] = 1.f; rmw.nroles = 1;
  SegLists sl;
  int rc = seg_build(rmw, valid, n, M, 0, ar, st, &sl);
  if (rc) return rc;
  const size_t sm4 = RescalGrouped<4>::smem_bytes(d), sm8 = RescalGrouped<8>::smem_bytes(d);
  const int vec16 = d % 4 == 0 && ((reinterpret_cast<uintptr_t>(E) | reinterpret_cast<uintptr_t>(W) |
                                    reinterpret_cast<uintptr_t>(G)) & 15) == 0;
  if (d <= RescalGrouped<4>::DMAX && sm4 <= 200 * 1024) {
    using RG = RescalGrouped<4>;
    // ask for more than half of the SM's shared memory: one CTA per SM, so that the ~n / 32 chunk CTAs of a
    // small minibatch spread over the SMs instead of pairing up
    const size_t smem = sm4 > 120 * 1024 ? sm4 : 120 * 1024;
    SKGE_CUDA(cudaFuncSetAttribute(rescal_logistic_grouped_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t chunks = (n + RG::EB - 1) / RG::EB;   // one CTA row per chunk (n < 2^27: seg_build)
    rescal_logistic_grouped_kernel<4><<<dim3((unsigned)chunks, RG::RUN_SPLIT), 256, smem, st>>>(
        E, W, s, o, p, y, sl.vals, sl.meta, d, vec16, G, fsv, loss, loss_accum, counts);
  } else if (d <= RescalGrouped<8>::DMAX && sm8 <= 200 * 1024) {
    using RG = RescalGrouped<8>;
    const size_t smem = sm8 > 120 * 1024 ? sm8 : 120 * 1024;
    SKGE_CUDA(cudaFuncSetAttribute(rescal_logistic_grouped_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t chunks = (n + RG::EB - 1) / RG::EB;
    rescal_logistic_grouped_kernel<8><<<dim3((unsigned)chunks, RG::RUN_SPLIT), 256, smem, st>>>(
        E, W, s, o, p, y, sl.vals, sl.meta, d, vec16, G, fsv, loss, loss_accum, counts);
  } else {   // W[p] does not fit in shared memory: one CTA per example, W streamed from L2
    int threads = block_threads(d);
    if (threads * 4 < d) threads = 256;
    SKGE_REQUIRE(threads * 4 >= d, "d too large");
    size_t smem = (3 * (size_t)d + 40) * sizeof(float);
    int64_t blocks = n > kNumSMs * 16 ? kNumSMs * 16 : n;
    rescal_logistic_kernel<<<(int)blocks, threads, smem, st>>>(E, W, s, o, p, y, valid, n, d, G, fsv, loss, loss_accum,
                                                               counts);
  }
  SKGE_LAUNCH_CHECK();
  // the per-relation outer-product mean (reads the OLD E)
  int tiles = (d + 31) / 32;
  int64_t items = uw * kGwSlices;
  dim3 grid(tiles, tiles, (unsigned)(items > 65535 ? 65535 : items));
  rescal_gw_partial_kernel<<<grid, 256, 0, st>>>(E, s, o, fsv, sl, d, gw_part, (int)uw);
  SKGE_LAUNCH_CHECK();
  int64_t dd = (int64_t)d * d;
  dim3 gridf((unsigned)((dd + 255) / 256 > 64 ? 64 : (dd + 255) / 256), (unsigned)(uw > 65535 ? 65535 : uw));
  rescal_gw_finish_kernel<<<gridf, 256, 0, st>>>(W, sl, d, rparam, gw_part, (int)uw, gw, pidx, counts);
  SKGE_LAUNCH_CHECK();
  // entity rows: keys ss+os get (fs*WE, fs*EW): rescal.py:72-74
  RoleMap rm;
  const int32_t *idx[2] = {s, o};
  for (int r = 0; r < 2; ++r) { rm.idx[r] = idx[r]; rm.is_rel[r] = 0; rm.grow[r] = r; rm.gsign[r] = 1.f; }
  rm.nroles = 2;
  ParamDesc pd[2];
  pd[0] = ParamDesc{E, p2E, postE, rparam, ucE, ge, eidx};
  pd[1] = ParamDesc{nullptr, nullptr, SKGE_POST_NONE, 0.f, nullptr, nullptr, nullptr};
  int32_t *counts_e = counts;  // seg_run writes counts[1] (U_E) and counts[2] (0 rows of table 1)
  // keep counts[2] = U_W: run the entity pass first into a scratch pair, then restore
  int32_t *scratch = ar.take<int32_t>(4);
  if (!ar.ok()) {
    set_error("workspace too small");
    return SKGE_EWORKSPACE;
  }
  rc = seg_run(rm, valid, n, N, 1, d, G, 2, pd, update, opt, lr, scratch, ar, st);
  if (rc) return rc;
  SKGE_CUDA(cudaMemcpyAsync(counts_e + 1, scratch + 1, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  if (update) {
    int64_t rowlen = (int64_t)d * d;
    int bx = (int)((rowlen + 255) / 256);
    if (bx > 64) bx = 64;
    dim3 g2(bx, (unsigned)(uw > 65535 ? 65535 : uw));
    rescal_w_update_kernel<<<g2, 256, 0, st>>>(W, p2W, gw, pidx, counts, rowlen, opt, lr, ucW);
    SKGE_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace skge

using namespace skge;

extern "C" {

size_t skge_logistic_workspace_bytes(int model, int64_t n, int d, int64_t N, int64_t M) {
  (void)N;
  return logistic_ws_bytes(model, n, d, M);
}

int skge_hole_logistic_grads(const float *E, const float *R, const int32_t *s, const int32_t *o,
                             const int32_t *p, const float *y, const uint8_t *valid, int64_t n, int64_t N,
                             int64_t M, int d, float rparam, float *ge, int32_t *eidx, float *gr,
                             int32_t *ridx, int32_t *counts, double *loss, void *ws,
                             size_t ws_bytes, skge_stream_t stream) {
  return hole_logistic_run(const_cast<float *>(E), const_cast<float *>(R), nullptr, nullptr, s, o, p, y, valid, n, N,
                           M, d, rparam, false, SKGE_OPT_SGD, 0.f, SKGE_POST_NONE, SKGE_POST_NONE, ge, eidx,
                           gr, ridx, counts, loss, nullptr, nullptr, nullptr, ws, ws_bytes, as_stream(stream));
}

int skge_hole_logistic_step(float *E, float *R, float *p2E, float *p2R, const int32_t *s,
                            const int32_t *o, const int32_t *p, const float *y, const uint8_t *valid,
                            int64_t n, int64_t N, int64_t M, int d, float rparam, int opt, float lr,
                            int postE, int postR, int32_t *counts, double *loss_accum,
                            int32_t *upd_counts_E, int32_t *upd_counts_R, void *ws,
                            size_t ws_bytes, skge_stream_t stream) {
  return hole_logistic_run(E, R, p2E, p2R, s, o, p, y, valid, n, N, M, d, rparam, true, opt, lr, postE, postR,
                           nullptr, nullptr, nullptr, nullptr, counts, nullptr, loss_accum, upd_counts_E,
                           upd_counts_R, ws, ws_bytes, as_stream(stream));
}

int skge_rescal_logistic_grads(const float *E, const float *W, const int32_t *s, const int32_t *o,
                               const int32_t *p, const float *y, const uint8_t *valid, int64_t n, int64_t N,
                               int64_t M, int d, float rparam, float *ge, int32_t *eidx, float *gw,
                               int32_t *pidx, int32_t *counts, double *loss, void *ws,
                               size_t ws_bytes, skge_stream_t stream) {
  return rescal_logistic_run(const_cast<float *>(E), const_cast<float *>(W), nullptr, nullptr, s, o, p, y, valid, n,
                             N, M, d, rparam, false, SKGE_OPT_SGD, 0.f, SKGE_POST_NONE, SKGE_POST_NONE, ge,
                             eidx, gw, pidx, counts, loss, nullptr, nullptr, nullptr, ws, ws_bytes,
                             as_stream(stream));
}

int skge_rescal_logistic_step(float *E, float *W, float *p2E, float *p2W, const int32_t *s,
                              const int32_t *o, const int32_t *p, const float *y, const uint8_t *valid,
                              int64_t n, int64_t N, int64_t M, int d, float rparam, int opt, float lr,
                              int postE, int postW, int32_t *counts, double *loss_accum,
                              int32_t *upd_counts_E, int32_t *upd_counts_W, void *ws,
                              size_t ws_bytes, skge_stream_t stream) {
  return rescal_logistic_run(E, W, p2E, p2W, s, o, p, y, valid, n, N, M, d, rparam, true, opt, lr, postE, postW,
                             nullptr, nullptr, nullptr, nullptr, counts, nullptr, loss_accum, upd_counts_E,
                             upd_counts_W, ws, ws_bytes, as_stream(stream));
}

}  // extern "C"
