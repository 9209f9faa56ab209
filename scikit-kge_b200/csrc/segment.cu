// Segment build (radix sort of row ids), segmented mean of gradient rows and the
// fused sparse SGD/AdaGrad update with row post-hooks.
//   grad_sum_matrix + Sm.dot(G)/n : skge/util.py:53-101, skge/transe.py:128-160,
//                                   skge/hole.py:31-40,69-97
//   SGD / AdaGrad / normalize / normless1 : skge/param.py:108-174
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <iterator>

#include "fft.cuh"
#include "segment.cuh"
#include "umma.cuh"

namespace skge {

// ---------------------------------------------------------------------------
// keys -> sorted segments
// ---------------------------------------------------------------------------

// One thread per unit: its role ids are read once and all of its (key, payload) entries written
// (entry t = role * P + unit, coalesced per role).
__global__ void build_keys_kernel(RoleMap rm, const uint8_t *__restrict__ flags, int64_t P, int N,
                                  int sentinel, int32_t *__restrict__ keys, int32_t *__restrict__ vals) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
    const bool live = flags ? flags[i] != 0 : true;
    int id[kMaxRoles];
#pragma unroll
    for (int r = 0; r < kMaxRoles; ++r) id[r] = r < rm.nroles ? rm.idx[r][i] : 0;
    // shared rows (see RoleMap::coef): roles (sp, sn, op, on, pp, pn)
    int shared = 0;
    if (rm.coef && id[4] == id[5]) {
      const bool same_s = id[0] == id[1], same_o = id[2] == id[3];
      shared = same_s && !same_o ? 1 : (same_o && !same_s ? 2 : 0);
    }
    const bool run_ok = rm.runw ? rm.runw[i] > 0 : true;
#pragma unroll
    for (int r = 0; r < kMaxRoles; ++r) {
      if (r >= rm.nroles) break;
      bool ok = live;
      const int q = rm.twin[r];
      bool same = false;
#pragma unroll
      for (int qq = 0; qq < kMaxRoles; ++qq)
        if (qq == q) same = id[qq] == id[r];
      if (same && q < r) ok = false;  // folded into the twin's row
      if (r == rm.runw_role && !run_ok) ok = false;  // folded into the head of its run
      int code = (same && q > r ? 8 : 0) + r;
      if (shared == 1 && (r == 2 || r == 3)) code = 6 + (r - 2);
      else if (shared == 2 && r < 2) code = 14 + r;
      keys[(int64_t)r * P + i] = ok ? id[r] + (rm.is_rel[r] ? N : 0) : sentinel;
      vals[(int64_t)r * P + i] = (int32_t)(i * 16 + code);
    }
  }
}

// head[i] = 1 where a segment starts in the sorted keys; evaluated on the fly by the scan and by the
// scatter pass (no pass of its own, no array)
struct HeadFlag {
  const int32_t *keys;
  int sentinel;
  __host__ __device__ __forceinline__ int32_t operator()(int64_t i) const {
    const int32_t k = keys[i];
    return (k != sentinel) && (i == 0 || k != keys[i - 1]);
  }
};
struct HeadIter {   // random-access input iterator over HeadFlag
  using value_type = int32_t;
  using difference_type = int64_t;
  using pointer = const int32_t *;
  using reference = int32_t;
  using iterator_category = std::random_access_iterator_tag;
  HeadFlag f;
  int64_t i;
  __host__ __device__ __forceinline__ int32_t operator*() const { return f(i); }
  __host__ __device__ __forceinline__ int32_t operator[](int64_t k) const { return f(i + k); }
  __host__ __device__ __forceinline__ HeadIter operator+(int64_t k) const { return HeadIter{f, i + k}; }
  __host__ __device__ __forceinline__ HeadIter operator-(int64_t k) const { return HeadIter{f, i - k}; }
  __host__ __device__ __forceinline__ int64_t operator-(const HeadIter &o) const { return i - o.i; }
  __host__ __device__ __forceinline__ HeadIter &operator+=(int64_t k) { i += k; return *this; }
  __host__ __device__ __forceinline__ HeadIter &operator++() { ++i; return *this; }
};

// meta: [0] nseg, [1] segments of table 0 (keys < N), [3] number of valid keys. Pre-zeroed.
__global__ void scatter_heads_kernel(const int32_t *__restrict__ keys,
                                     const int32_t *__restrict__ pos, int64_t L, int N, int sentinel,
                                     int32_t *__restrict__ seg_start, int32_t *__restrict__ seg_key,
                                     int32_t *__restrict__ meta) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < L; i += (int64_t)gridDim.x * blockDim.x) {
    int k = keys[i];
    if (k == sentinel) continue;
    int h = i == 0 || keys[i - 1] != k, ps = pos[i];
    if (h) {
      seg_start[ps] = (int32_t)i;
      seg_key[ps] = k;
      if (k >= N && (i == 0 || keys[i - 1] < N)) meta[1] = ps;  // first segment of table 1
    }
    if (i == L - 1 || keys[i + 1] == sentinel) {  // last valid key
      int nseg = ps + h;
      meta[0] = nseg;
      meta[3] = (int32_t)(i + 1);
      seg_start[nseg] = (int32_t)(i + 1);
      if (k < N) meta[1] = nseg;  // no table-1 segments at all
    }
  }
}

__global__ void set_i32_kernel(int32_t *p, int32_t v) { *p = v; }
int set_i32(int32_t *p, int32_t v, cudaStream_t st) {
  set_i32_kernel<<<1, 1, 0, st>>>(p, v);
  SKGE_LAUNCH_CHECK();
  return 0;
}

static int key_bits(int64_t maxkey) {
  int b = 1;
  while (((int64_t)1 << b) <= maxkey) ++b;
  return b;
}

static size_t cub_sort_bytes(int64_t L) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const int32_t *)nullptr, (int32_t *)nullptr,
                                  (const int32_t *)nullptr, (int32_t *)nullptr, (int)L, 0, 32);
  return bytes;
}
__global__ void iota_kernel(int32_t *v, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) v[i] = (int32_t)i;
}
size_t order_workspace_bytes(int64_t n) {
  if (n < 1) n = 1;
  return 3 * align_up((size_t)n * 4) + align_up(cub_sort_bytes(n)) + 256;
}
int order_by_key(const int32_t *keys, int64_t n, int key_bits, Arena &ar, cudaStream_t st, const int32_t **order) {
  SKGE_REQUIRE(n > 0 && n < ((int64_t)1 << 31), "bad length");
  int32_t *keys_out = ar.take<int32_t>(n), *vals_in = ar.take<int32_t>(n), *vals_out = ar.take<int32_t>(n);
  size_t tb = cub_sort_bytes(n);
  void *tmp = ar.take<char>(tb);
  if (!ar.ok()) {
    set_error("workspace too small: need %zu bytes, have %zu", ar.off, ar.cap);
    return SKGE_EWORKSPACE;
  }
  int blocks = (int)((n + 255) / 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  iota_kernel<<<blocks, 256, 0, st>>>(vals_in, n);
  SKGE_LAUNCH_CHECK();
  SKGE_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, keys, keys_out, vals_in, vals_out, (int)n, 0, key_bits, st));
  *order = vals_out;
  return 0;
}

static size_t cub_scan_bytes(int64_t L) {
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, HeadIter{HeadFlag{nullptr, 0}, 0}, (int32_t *)nullptr, (int)L);
  return bytes;
}

// Segments longer than seg_chunk occurrences (hot rows: a hub entity, a frequent relation)
// are cut into chunks reduced by different warps and combined by one CTA per segment, in a
// fixed order (deterministic, no atomics on parameter rows).
// The chunk length trades parallelism against partial-sum traffic: small minibatches (latency
// bound: the longest serial chain sets the kernel time) use 32, large ones 128.
static int seg_chunk_for(int64_t L) { return L <= (1 << 18) ? 32 : 128; }
static int64_t long_chunk_cap(int64_t L) { return L / (seg_chunk_for(L) / 2) + 8; }  // sum ceil(len/c) over len > c
static int64_t long_seg_cap(int64_t L) { return L / seg_chunk_for(L) + 8; }

size_t seg_workspace_bytes(int64_t L, int d) {
  if (L < 1) L = 1;
  size_t b = 0;
  b += align_up((size_t)long_seg_cap(L) * 3 * 4) + align_up((size_t)long_chunk_cap(L) * 3 * 4) + 256 + align_up(16);
  b += align_up((size_t)long_chunk_cap(L) * (d > 0 ? d : 1) * sizeof(float));
  b += 4 * align_up((size_t)L * 4);        // keys in/out, vals in/out
  b += 2 * align_up((size_t)L * 4);        // head, pos
  b += 2 * align_up((size_t)(L + 1) * 4);  // seg_start, seg_key
  b += align_up(16);                       // meta
  size_t c1 = cub_sort_bytes(L), c2 = cub_scan_bytes(L);
  b += align_up(c1 > c2 ? c1 : c2);
  return b + 1024;
}

int seg_build(const RoleMap &rm, const uint8_t *flags, int64_t P, int64_t N, int64_t M, Arena &ar,
              cudaStream_t st, SegLists *out) {
  int64_t L = (int64_t)rm.nroles * P;
  SKGE_REQUIRE(L > 0 && L < ((int64_t)1 << 31) && P < ((int64_t)1 << 27), "minibatch too large");
  SKGE_REQUIRE(N + M < ((int64_t)1 << 31) - 1, "too many rows");
  int32_t *keys_in = ar.take<int32_t>(L), *keys_out = ar.take<int32_t>(L);
  int32_t *vals_in = ar.take<int32_t>(L), *vals_out = ar.take<int32_t>(L);
  int32_t *pos = ar.take<int32_t>(L);
  int32_t *seg_start = ar.take<int32_t>(L + 1), *seg_key = ar.take<int32_t>(L + 1);
  int32_t *meta = ar.take<int32_t>(4);
  size_t c1 = cub_sort_bytes(L), c2 = cub_scan_bytes(L);
  size_t cub_bytes = c1 > c2 ? c1 : c2;
  void *cub_tmp = ar.take<char>(cub_bytes);
  if (!ar.ok()) {
    set_error("workspace too small: need %zu bytes, have %zu", ar.off, ar.cap);
    return SKGE_EWORKSPACE;
  }
  int sentinel = (int)(N + M);
  int threads = 256;
  int blocks = (int)((L + threads - 1) / threads);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  SKGE_CUDA(cudaMemsetAsync(meta, 0, 16, st));
  {
    int kb = (int)((P + threads - 1) / threads);
    if (kb > kNumSMs * 8) kb = kNumSMs * 8;
    build_keys_kernel<<<kb, threads, 0, st>>>(rm, flags, P, (int)N, sentinel, keys_in, vals_in);
  }
  SKGE_LAUNCH_CHECK();
  size_t tb = cub_bytes;
  SKGE_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, tb, keys_in, keys_out, vals_in, vals_out, (int)L, 0,
                                            key_bits(sentinel), st));
  tb = cub_bytes;
  SKGE_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp, tb, HeadIter{HeadFlag{keys_out, sentinel}, 0}, pos, (int)L, st));
  scatter_heads_kernel<<<blocks, threads, 0, st>>>(keys_out, pos, L, (int)N, sentinel, seg_start, seg_key, meta);
  SKGE_LAUNCH_CHECK();
  out->vals = vals_out;
  out->seg_start = seg_start;
  out->seg_key = seg_key;
  out->meta = meta;
  return 0;
}

// ---------------------------------------------------------------------------
// row update (shared by the fused segment kernel and skge_sparse_update)
// ---------------------------------------------------------------------------

// One warp owns one row of d floats, held as g[MAXC][VEC] per lane
// (column = (c*32 + lane)*VEC + v).  Applies rparam, the optimiser step and the
// post-hook and writes the row (and p2) exactly once.
template <int VEC, int MAXC>
__device__ __forceinline__ void row_update(float *xrow, float *p2row, float (&g)[MAXC][VEC], int d,
                                           int lane, int opt, float lr, int post, float rparam,
                                           float (&x)[MAXC][VEC]);

template <int VEC, int MAXC>
__device__ __forceinline__ void row_update(float *xrow, float *p2row, float (&g)[MAXC][VEC], int d,
                                           int lane, int opt, float lr, int post, float rparam) {
  float x[MAXC][VEC];
  row_update<VEC, MAXC>(xrow, p2row, g, d, lane, opt, lr, post, rparam, x);
}

template <int VEC, int MAXC>
__device__ __forceinline__ void row_update(float *xrow, float *p2row, float (&g)[MAXC][VEC], int d,
                                           int lane, int opt, float lr, int post, float rparam,
                                           float (&x)[MAXC][VEC]) {
  float ss = 0.f;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    int col = (c * 32 + lane) * VEC;
    if (col < d) {
      ld_vec_rw<VEC>(xrow + col, x[c]);
      if (opt == SKGE_OPT_ADAGRAD) {
        float p2[VEC];
        ld_vec_rw<VEC>(p2row + col, p2);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          float gg = g[c][v] + rparam * x[c][v];
          p2[v] += gg * gg;                                   // skge/param.py:147
          x[c][v] -= lr * gg * adagrad_rscale(p2[v]);         // skge/param.py:152-155
        }
        st_vec<VEC>(p2row + col, p2);
      } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) x[c][v] -= lr * (g[c][v] + rparam * x[c][v]);  // skge/param.py:130
      }
#pragma unroll
      for (int v = 0; v < VEC; ++v) ss += x[c][v] * x[c][v];
    }
  }
  float scale = 1.f;
  if (post != SKGE_POST_NONE) {
    ss = warp_sum(ss);
    if (post == SKGE_POST_NORMALIZE) scale = rsqrtf(ss);                 // skge/param.py:165-166
    else scale = 1.0f / (ss < 1.0f ? 1.0f : ss);                        // skge/param.py:171-173 (squared norm)
  }
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    int col = (c * 32 + lane) * VEC;
    if (col < d) {
      if (post != SKGE_POST_NONE) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) x[c][v] *= scale;
      }
      st_vec<VEC>(xrow + col, x[c]);
    }
  }
}

struct SegArgs {
  const int32_t *seg_start, *seg_key, *vals, *meta;
  const float *G;
  int rows_per_unit, d, N;
  int grow[8];
  float gsign[8];
  ParamDesc pd[2];
  int opt;
  float lr;
  int32_t *counts;
  // long segments (more than seg_chunk occurrences)
  int32_t *long_meta;   // [0] number of long segments, [1] number of chunks allocated
  int32_t *long_seg;    // [long_seg_cap][3]: segment, first chunk, number of chunks
  int32_t *long_work;   // [long_chunk_cap][2]: segment, chunk index inside the segment
  float *partials;      // [long_chunk_cap][d]
  int32_t *partial_occ; // [long_chunk_cap] occurrences behind each partial sum
  int long_seg_cap, long_chunk_cap;
  int seg_chunk;        // segments longer than this are reduced chunk-wise
  int spec_d;           // > 0: G rows are packed spectra of length spec_d = d (see fft.cuh)
  const int32_t *runw;  // RoleMap::runw (nullable)
  int runw_role;
  const float *coef;    // RoleMap::coef (nullable)
};

// What a sorted payload (unit * 16 + code) refers to: the gradient row (index into G, in rows), the
// factor it is taken with, and the occurrences it stands for in the mean: a run sum (RoleMap::runw),
// a folded twin row (2) or 1.  Codes 6, 7, 14, 15 are the shared rows of RoleMap::coef.
struct OccRef {
  int goff;
  float coef;
  int weight;
};
__device__ __forceinline__ OccRef occ_decode(const SegArgs &a, int val) {
  const int c = val & 15, unit = val >> 4;
  OccRef o;
  if (a.coef && (c & 6) == 6) {
    o.goff = unit * a.rows_per_unit + (c < 8 ? 2 : 0);
    o.coef = a.coef[2 * unit + (c & 1)];
    o.weight = 1;
  } else {
    const int r = c & 7;
    o.goff = unit * a.rows_per_unit + a.grow[r];
    o.coef = a.gsign[r];
    o.weight = (a.runw && r == a.runw_role) ? a.runw[unit] : 1 + (c >> 3);
  }
  return o;
}

// shared memory of the spectral mode: twiddles [d/2] float2, then per warp two complex
// buffers of d/2 float2 (2 d floats)
// (d = 256 takes the register-resident transforms of fft.cuh: 1 KB of transposition space per warp)
__host__ __device__ __forceinline__ size_t spec_smem_bytes(int d, int warps) {
  if (d == 256) return (size_t)warps * 1024;
  return (size_t)(d / 2) * sizeof(float2) + (size_t)warps * warp_fft_scratch_floats(d) * sizeof(float);
}
template <int VEC, int MAXC, bool UPDATE, bool SPEC>
struct SpecReg { static constexpr bool value = UPDATE && SPEC && VEC == 4 && MAXC == 2; };   // <=> d == 256

// acc += signed gradient rows of occurrences [beg, end) of the sorted list.  The payloads of
// up to 32 occurrences are fetched with one coalesced load and broadcast by shuffle; row loads
// are issued BATCH at a time before they are consumed (memory-level parallelism: a warp that
// walks a long segment is latency bound otherwise).  BATCH = 4 for small, latency-bound
// minibatches; BATCH = 1 keeps the register count (hence the occupancy) up for the large,
// bandwidth-bound ones.
// Returns the number of occurrences the rows stand for (a folded twin row counts twice).
template <int VEC, int MAXC, int BATCH>
__device__ __forceinline__ int accumulate_rows(const SegArgs &a, int beg, int end, int lane,
                                               float (&acc)[MAXC][VEC]) {
  const int d = a.d;
  int occ = 0;
  for (int j0 = beg; j0 < end; j0 += 32) {
    const int cnt = min(32, end - j0);
    OccRef mine = {0, 0.f, 0};
    if (lane < cnt) mine = occ_decode(a, a.vals[j0 + lane]);
    occ += __reduce_add_sync(kFull, mine.weight);
    for (int t0 = 0; t0 < cnt; t0 += BATCH) {
      float tmp[BATCH][MAXC][VEC];
      float sgn[BATCH];
#pragma unroll
      for (int b = 0; b < BATCH; ++b) {
        const int src = min(t0 + b, cnt - 1);
        const int goff = __shfl_sync(kFull, mine.goff, src);
        const float cf = __shfl_sync(kFull, mine.coef, src);
        const bool live = t0 + b < cnt;
        sgn[b] = live ? cf : 0.f;
        const float *g = a.G + (int64_t)goff * d;
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
          int col = (c * 32 + lane) * VEC;
          if (col < d) ld_vec<VEC>(g + col, tmp[b][c]);
        }
      }
#pragma unroll
      for (int b = 0; b < BATCH; ++b)
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
          int col = (c * 32 + lane) * VEC;
          if (col < d) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[c][v] = fmaf(sgn[b], tmp[b][c][v], acc[c][v]);
          }
        }
    }
  }
  return occ;
}

// acc holds the SUM over the n occurrences of segment seg: take the mean and either apply
// the optimiser step in place or emit (gradient row, row id).
// d = 256, spectral update: acc is the summed packed spectrum of the row's n occurrences.  Inverse
// transform, mean, optimiser step and post-hook in the time domain, forward transform of the updated
// row into the spectral table -- all in registers (fft.cuh), two shared-memory transpositions.
__device__ __forceinline__ void finish_row_spec256(const SegArgs &a, int key, int n, int lane, float (&acc)[2][4],
                                                   float *spec_smem, int warp_in_cta, const RegFft256 &rc) {
  constexpr int d = 256;
  const int which = key >= a.N;
  const int64_t row = which ? key - a.N : key;
  const ParamDesc &pd = a.pd[which];
  float4 *buf = reinterpret_cast<float4 *>(spec_smem) + warp_in_cta * 64;
  float2 v[4];
  regfft256_row_to_freq(buf, make_float4(acc[0][0], acc[0][1], acc[0][2], acc[0][3]),
                        make_float4(acc[1][0], acc[1][1], acc[1][2], acc[1][3]), v, lane);
  regfft256_irfft(v, rc, lane);
  const float sc = (2.0f / (float)d) / (float)n;   // transform scale and the mean (skge/util.py:97-101)
  float g[4][2], x[4][2];
#pragma unroll
  for (int j = 0; j < 4; ++j) { g[j][0] = v[j].x * sc; g[j][1] = v[j].y * sc; }
  row_update<2, 4>(pd.param + row * d, pd.p2 ? pd.p2 + row * d : nullptr, g, d, lane, a.opt, a.lr, pd.post, pd.rparam, x);
  if (pd.upd_counts && lane == 0) pd.upd_counts[row] += 1;
  if (pd.hat) {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = make_float2(x[j][0], x[j][1]);
    regfft256_rfft(v, rc, lane);
    float4 r0, r1;
    regfft256_freq_to_row(buf, v, r0, r1, lane);
    float4 *hrow = reinterpret_cast<float4 *>(pd.hat + row * d);
    hrow[lane] = r0;
    hrow[32 + lane] = r1;
  }
}

template <int VEC, int MAXC, bool UPDATE, bool SPEC>
__device__ __forceinline__ void finish_row(const SegArgs &a, int seg, int key, int n, int lane,
                                           float (&acc)[MAXC][VEC], float *spec_smem, int warp_in_cta,
                                           const RegFft256 &rc) {
  if constexpr (SpecReg<VEC, MAXC, UPDATE, SPEC>::value) {
    finish_row_spec256(a, key, n, lane, acc, spec_smem, warp_in_cta, rc);
    return;
  }
  const int d = a.d;
  const int U0 = a.meta[1];
  int which = key >= a.N;
  int64_t row = which ? key - a.N : key;
  const float inv_n = 1.0f / (float)n;
  float2 *tw = nullptr, *b0 = nullptr, *b1 = nullptr;
  if (UPDATE && SPEC) {
    // acc is a summed packed spectrum: back to the time domain
    const int h = d / 2;
    tw = reinterpret_cast<float2 *>(spec_smem);
    b0 = tw + h + (size_t)warp_in_cta * 2 * h;
    b1 = b0 + h;
    float *pk = reinterpret_cast<float *>(b1);   // the packed row may live in b1 until the first stage
    __syncwarp();
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      int col = (c * 32 + lane) * VEC;
      if (col < d) st_vec<VEC>(pk + col, acc[c]);
    }
    const float *x = warp_irfft_packed(pk, b0, b1, tw, a.spec_d, lane);
    const float sc = 2.0f / (float)d;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      int col = (c * 32 + lane) * VEC;
      if (col < d) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[c][v] = x[col + v] * sc;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < MAXC; ++c)
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[c][v] *= inv_n;  // the mean: skge/util.py:97-101
  const ParamDesc &pd = a.pd[which];
  if (UPDATE) {
    float x[MAXC][VEC];
    row_update<VEC, MAXC>(pd.param + row * d, pd.p2 ? pd.p2 + row * d : nullptr, acc, d, lane, a.opt, a.lr,
                          pd.post, pd.rparam, x);
    if (pd.upd_counts && lane == 0) pd.upd_counts[row] += 1;
    if (SPEC && pd.hat) {
      // refresh the row's packed spectrum from the updated time-domain row
      const int h = d / 2;
      __syncwarp();
      float *xr = reinterpret_cast<float *>(b0);
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        int col = (c * 32 + lane) * VEC;
        if (col < d) st_vec<VEC>(xr + col, x[c]);
      }
      const float2 *Z = warp_rfft_half(b0, b1, tw, a.spec_d, lane);
      float2 *hrow = reinterpret_cast<float2 *>(pd.hat + row * d);
      for (int f = lane; f < h; f += 32) hrow[f] = packed_slot(Z, f, h, tw);
    }
  } else {
    int64_t u = which ? seg - U0 : seg;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      int col = (c * 32 + lane) * VEC;
      if (col < d) {
        if (pd.rparam != 0.f) {
          float x[VEC];
          ld_vec<VEC>(pd.param + row * d + col, x);
#pragma unroll
          for (int v = 0; v < VEC; ++v) acc[c][v] += pd.rparam * x[v];
        }
        st_vec<VEC>(pd.out_g + u * d + col, acc[c]);
      }
    }
    if (lane == 0) pd.out_idx[u] = (int32_t)row;
  }
}

// Pass 1: one warp per segment.  Short segments are finished here; long ones are registered
// (segment, chunk range) for passes 2 and 3.
#ifndef SKGE_SEG_SPEC256_CTAS
#define SKGE_SEG_SPEC256_CTAS 4
#endif
template <int VEC, int MAXC, bool UPDATE, int BATCH, bool SPEC>
__global__ void __launch_bounds__(256, SpecReg<VEC, MAXC, UPDATE, SPEC>::value ? SKGE_SEG_SPEC256_CTAS
                                       : ((BATCH > 1 || MAXC > 2) ? 1 : 4)) seg_reduce_kernel(SegArgs a) {
  extern __shared__ __align__(16) float spec_smem[];
  const int lane = threadIdx.x & 31;
  RegFft256 rc;
  if constexpr (SpecReg<VEC, MAXC, UPDATE, SPEC>::value) {
    regfft256_init(rc, lane);
  } else if (UPDATE && SPEC) {
    fill_twiddles(reinterpret_cast<float2 *>(spec_smem), a.d, threadIdx.x, blockDim.x);
    __syncthreads();
  }
  const int nseg = a.meta[0];
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.counts) {
    a.counts[1] = a.meta[1];
    a.counts[2] = nseg - a.meta[1];
  }
  int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int nwarps = gridDim.x * (blockDim.x >> 5);
  for (int seg = warp; seg < nseg; seg += nwarps) {
    int key = a.seg_key[seg];
    int beg = a.seg_start[seg], end = a.seg_start[seg + 1];
    int n = end - beg;
    if (n > a.seg_chunk) {
      int nch = (n + a.seg_chunk - 1) / a.seg_chunk;
      int slot = 0, first = 0;
      if (lane == 0) {
        slot = atomicAdd(a.long_meta, 1);
        first = atomicAdd(a.long_meta + 1, nch);
      }
      slot = __shfl_sync(kFull, slot, 0);
      first = __shfl_sync(kFull, first, 0);
      if (lane == 0) {
        a.long_seg[3 * slot] = seg;
        a.long_seg[3 * slot + 1] = first;
        a.long_seg[3 * slot + 2] = nch;
      }
      for (int k = lane; k < nch; k += 32) {
        a.long_work[2 * (first + k)] = seg;
        a.long_work[2 * (first + k) + 1] = k;
      }
      continue;
    }
    float acc[MAXC][VEC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[c][v] = 0.f;
    const int occ = accumulate_rows<VEC, MAXC, BATCH>(a, beg, end, lane, acc);
    finish_row<VEC, MAXC, UPDATE, SPEC>(a, seg, key, occ, lane, acc, spec_smem, threadIdx.x >> 5, rc);
  }
}

// ---------------------------------------------------------------------------
// Pass 1 for large minibatches of 1 KB rows (d = 256), update mode: the same walk, but every row a
// segment needs -- its parameter row, its AdaGrad row and up to SLOTS - 2 gradient rows at a time --
// is brought into a per-warp shared-memory ring by 1-D bulk-TMA copies (one row = one contiguous
// 1 KB copy, issued by one lane each, all completing on the warp's own mbarrier).  The register
// walk above keeps ONE row (1 KB) in flight per warp, and with ~1.2 us of loaded DRAM latency the
// SM then idles on memory (ncu: 44 % of DRAM peak at 28 resident warps); here a warp has
// (n + 2) KB in flight without spending a register on it.
// ---------------------------------------------------------------------------
template <int VEC, int MAXC>
__device__ __forceinline__ void row_update_staged(float *xrow, float *p2row, float (&g)[MAXC][VEC],
                                                  float (&x)[MAXC][VEC], float (&p2)[MAXC][VEC], int lane, int opt,
                                                  float lr, int post, float rparam) {
  float ss = 0.f;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    if (opt == SKGE_OPT_ADAGRAD) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        float gg = g[c][v] + rparam * x[c][v];
        p2[c][v] += gg * gg;                                   // skge/param.py:147
        x[c][v] -= lr * gg * adagrad_rscale(p2[c][v]);         // skge/param.py:152-155
      }
      st_vec<VEC>(p2row + (c * 32 + lane) * VEC, p2[c]);
    } else {
#pragma unroll
      for (int v = 0; v < VEC; ++v) x[c][v] -= lr * (g[c][v] + rparam * x[c][v]);  // skge/param.py:130
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) ss += x[c][v] * x[c][v];
  }
  float scale = 1.f;
  if (post != SKGE_POST_NONE) {
    ss = warp_sum(ss);
    if (post == SKGE_POST_NORMALIZE) scale = rsqrtf(ss);                 // skge/param.py:165-166
    else scale = 1.0f / (ss < 1.0f ? 1.0f : ss);                        // skge/param.py:171-173 (squared norm)
  }
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    if (post != SKGE_POST_NONE) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) x[c][v] *= scale;
    }
    st_vec<VEC>(xrow + (c * 32 + lane) * VEC, x[c]);
  }
}

#ifndef SKGE_SEG_BULK_SLOTS
#define SKGE_SEG_BULK_SLOTS 8
#endif
#ifndef SKGE_SEG_BULK_CTAS
#define SKGE_SEG_BULK_CTAS 3
#endif
static constexpr int kBulkRowFloats = 256;
static constexpr int kBulkRowBytes = kBulkRowFloats * 4;
__host__ __device__ constexpr size_t seg_bulk_smem_bytes(int slots) { return 128 + (size_t)8 * slots * kBulkRowBytes; }

template <bool SPEC, int SLOTS>
__global__ void __launch_bounds__(256, SKGE_SEG_BULK_CTAS) seg_reduce_bulk_kernel(SegArgs a) {
  extern __shared__ __align__(128) unsigned char bulk_smem[];
  constexpr int d = kBulkRowFloats, CAP = SLOTS - 2;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint64_t *bar = reinterpret_cast<uint64_t *>(bulk_smem) + w;
  float4 *slots = reinterpret_cast<float4 *>(bulk_smem + 128) + (size_t)w * SLOTS * (d / 4);   // [SLOTS][64] float4
  if (lane == 0) {
    ptx::mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  RegFft256 rc;
  if constexpr (SPEC) regfft256_init(rc, lane);
  const int nseg = a.meta[0];
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.counts) {
    a.counts[1] = a.meta[1];
    a.counts[2] = nseg - a.meta[1];
  }
  const bool adagrad = a.opt == SKGE_OPT_ADAGRAD;
  uint32_t phase = 0;
  // A warp takes 32 consecutive segments at a time (dealt out through a counter: the relation rows,
  // whose runs are the longest, sit at the end of the key order and are started first).  The segment
  // table entries and a 32-wide window of decoded occurrences live one per lane and are handed
  // round by shuffles, so the dependent chain table -> payload -> (gp, gn) / run weight -> row address
  // is paid once per 32 segments / occurrences instead of once per segment.
  const int nblk = (nseg + 31) >> 5;
  for (;;) {
    int blk = 0;
    if (lane == 0) blk = atomicAdd(a.long_meta + 2, 1);
    blk = __shfl_sync(kFull, blk, 0);
    if (blk >= nblk) break;
    const int s0 = (nblk - 1 - blk) << 5;
    const int ns = min(32, nseg - s0);
    int mkey = 0, mbeg = 0, mend = 0;
    if (lane < ns) {
      mkey = a.seg_key[s0 + lane];
      mbeg = a.seg_start[s0 + lane];
      mend = a.seg_start[s0 + lane + 1];
    }
    const int wend = __shfl_sync(kFull, mend, ns - 1);
    int w0 = 0;
    OccRef mine;
    auto load_window = [&](int base) {
      w0 = base;
      mine = OccRef{0, 0.f, 0};
      if (base + lane < wend) mine = occ_decode(a, a.vals[base + lane]);
    };
    load_window(__shfl_sync(kFull, mbeg, 0));
  for (int sj = 0; sj < ns; ++sj) {
    const int seg = s0 + sj;
    const int key = __shfl_sync(kFull, mkey, sj);
    const int beg = __shfl_sync(kFull, mbeg, sj), end = __shfl_sync(kFull, mend, sj);
    const int n = end - beg;
    if (n > a.seg_chunk) {   // hot row: passes 2 and 3
      int nch = (n + a.seg_chunk - 1) / a.seg_chunk;
      int slot = 0, first = 0;
      if (lane == 0) {
        slot = atomicAdd(a.long_meta, 1);
        first = atomicAdd(a.long_meta + 1, nch);
      }
      slot = __shfl_sync(kFull, slot, 0);
      first = __shfl_sync(kFull, first, 0);
      if (lane == 0) {
        a.long_seg[3 * slot] = seg;
        a.long_seg[3 * slot + 1] = first;
        a.long_seg[3 * slot + 2] = nch;
      }
      for (int k = lane; k < nch; k += 32) {
        a.long_work[2 * (first + k)] = seg;
        a.long_work[2 * (first + k) + 1] = k;
      }
      continue;
    }
    const int which = key >= a.N;
    const int64_t row = which ? key - a.N : key;
    const ParamDesc &pd = a.pd[which];
    float *xrow = pd.param + row * d, *p2row = adagrad ? pd.p2 + row * d : nullptr;
    float acc[2][4];
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int v = 0; v < 4; ++v) acc[c][v] = 0.f;
    int occ = 0;
    for (int j0 = beg; j0 < end; j0 += CAP) {
      const int cnt = min(CAP, end - j0);
      if (j0 + cnt > w0 + 32) load_window(j0);
      const int src = (j0 - w0 + lane) & 31;   // lane t < cnt: where occurrence j0 + t sits in the window
      const int goff = __shfl_sync(kFull, mine.goff, src);
      const int wgt = __shfl_sync(kFull, mine.weight, src);
      occ += __reduce_add_sync(kFull, lane < cnt ? wgt : 0);
      const bool first = j0 == beg;
      // the slots were last touched by this warp's own (generic-proxy) reads and writes
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) ptx::mbar_expect_tx(bar, (uint32_t)(cnt + (first ? (adagrad ? 2 : 1) : 0)) * kBulkRowBytes);
      __syncwarp();
      if (lane < cnt)
        ptx::bulk_g2s(slots + (2 + lane) * (d / 4), a.G + (int64_t)goff * d, kBulkRowBytes, bar);
      if (first) {
        if (lane == 30) ptx::bulk_g2s(slots, xrow, kBulkRowBytes, bar);
        if (lane == 31 && adagrad) ptx::bulk_g2s(slots + d / 4, p2row, kBulkRowBytes, bar);
      }
      ptx::mbar_wait(bar, phase);
      phase ^= 1;
      for (int t = 0; t < cnt; ++t) {
        const float sgn = __shfl_sync(kFull, mine.coef, j0 - w0 + t);
        const float4 u0 = slots[(2 + t) * (d / 4) + lane], u1 = slots[(2 + t) * (d / 4) + 32 + lane];
        acc[0][0] = fmaf(sgn, u0.x, acc[0][0]); acc[0][1] = fmaf(sgn, u0.y, acc[0][1]);
        acc[0][2] = fmaf(sgn, u0.z, acc[0][2]); acc[0][3] = fmaf(sgn, u0.w, acc[0][3]);
        acc[1][0] = fmaf(sgn, u1.x, acc[1][0]); acc[1][1] = fmaf(sgn, u1.y, acc[1][1]);
        acc[1][2] = fmaf(sgn, u1.z, acc[1][2]); acc[1][3] = fmaf(sgn, u1.w, acc[1][3]);
      }
    }
    if constexpr (SPEC) {
      // summed packed spectrum -> time domain (registers), update, -> spectrum of the updated row
      float2 v[4];
      float4 *tbuf = slots + 2 * (d / 4);   // a consumed gradient slot serves the two transpositions
      regfft256_row_to_freq(tbuf, make_float4(acc[0][0], acc[0][1], acc[0][2], acc[0][3]),
                            make_float4(acc[1][0], acc[1][1], acc[1][2], acc[1][3]), v, lane);
      regfft256_irfft(v, rc, lane);
      const float sc = (2.0f / (float)d) / (float)occ;   // transform scale and the mean (skge/util.py:97-101)
      float g[4][2], x[4][2], p2[4][2];
      const float2 *xs = reinterpret_cast<const float2 *>(slots), *ps = reinterpret_cast<const float2 *>(slots + d / 4);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        g[j][0] = v[j].x * sc; g[j][1] = v[j].y * sc;
        const float2 xv = xs[32 * j + lane];
        x[j][0] = xv.x; x[j][1] = xv.y;
        if (adagrad) { const float2 pv = ps[32 * j + lane]; p2[j][0] = pv.x; p2[j][1] = pv.y; }
      }
      row_update_staged<2, 4>(xrow, p2row, g, x, p2, lane, a.opt, a.lr, pd.post, pd.rparam);
      if (pd.hat) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = make_float2(x[j][0], x[j][1]);
        regfft256_rfft(v, rc, lane);
        float4 r0, r1;
        regfft256_freq_to_row(tbuf, v, r0, r1, lane);
        float4 *hrow = reinterpret_cast<float4 *>(pd.hat + row * d);
        hrow[lane] = r0;
        hrow[32 + lane] = r1;
      }
    } else {
      const float inv_n = 1.0f / (float)occ;
      float x[2][4], p2[2][4];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[c][v] *= inv_n;   // the mean: skge/util.py:97-101
        const float4 xv = slots[c * 32 + lane];
        x[c][0] = xv.x; x[c][1] = xv.y; x[c][2] = xv.z; x[c][3] = xv.w;
        if (adagrad) {
          const float4 pv = slots[d / 4 + c * 32 + lane];
          p2[c][0] = pv.x; p2[c][1] = pv.y; p2[c][2] = pv.z; p2[c][3] = pv.w;
        }
      }
      row_update_staged<4, 2>(xrow, p2row, acc, x, p2, lane, a.opt, a.lr, pd.post, pd.rparam);
    }
    if (pd.upd_counts && lane == 0) pd.upd_counts[row] += 1;
  }
  }
}

static bool seg_bulk_ok(const SegArgs &a) {
  if (a.d != kBulkRowFloats || a.seg_chunk <= 32) return false;
  auto al = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  for (int t = 0; t < 2; ++t)
    if (!al(a.pd[t].param) || !al(a.pd[t].p2) || !al(a.pd[t].hat)) return false;
  return al(a.G);
}

// Pass 2: one warp per chunk of a long segment -> partial sum row.
template <int VEC, int MAXC, int BATCH>
__global__ void __launch_bounds__(256) seg_long_chunks_kernel(SegArgs a) {
  const int lane = threadIdx.x & 31;
  const int nchunks = a.long_meta[1];
  int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int nwarps = gridDim.x * (blockDim.x >> 5);
  const int d = a.d;
  for (int c = warp; c < nchunks; c += nwarps) {
    int seg = a.long_work[2 * c], k = a.long_work[2 * c + 1];
    int beg = a.seg_start[seg] + k * a.seg_chunk;
    int end = min(a.seg_start[seg + 1], beg + a.seg_chunk);
    float acc[MAXC][VEC];
#pragma unroll
    for (int cc = 0; cc < MAXC; ++cc)
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[cc][v] = 0.f;
    const int occ = accumulate_rows<VEC, MAXC, BATCH>(a, beg, end, lane, acc);
    if (lane == 0) a.partial_occ[c] = occ;
    float *dst = a.partials + (int64_t)c * d;
#pragma unroll
    for (int cc = 0; cc < MAXC; ++cc) {
      int col = (cc * 32 + lane) * VEC;
      if (col < d) st_vec<VEC>(dst + col, acc[cc]);
    }
  }
}

// Pass 3: one CTA per long segment: warp w sums partials w, w + 8, ...; the eight warp sums are
// combined through shared memory in warp order, then warp 0 finishes the row.
template <int VEC, int MAXC, bool UPDATE, bool SPEC>
__global__ void __launch_bounds__(256) seg_long_finish_kernel(SegArgs a) {
  extern __shared__ __align__(16) float red[];  // [8][d], then the spectral scratch
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int nlong = a.long_meta[0];
  const int d = a.d;
  float *spec_smem = red + 8 * d;
  RegFft256 rc;
  if constexpr (SpecReg<VEC, MAXC, UPDATE, SPEC>::value) {
    regfft256_init(rc, lane);
  } else if (UPDATE && SPEC) {
    fill_twiddles(reinterpret_cast<float2 *>(spec_smem), d, threadIdx.x, blockDim.x);
    __syncthreads();
  }
  for (int s = blockIdx.x; s < nlong; s += gridDim.x) {
    int seg = a.long_seg[3 * s], first = a.long_seg[3 * s + 1], nch = a.long_seg[3 * s + 2];
    float acc[MAXC][VEC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[c][v] = 0.f;
    for (int k = w; k < nch; k += 8) {
      const float *src = a.partials + (int64_t)(first + k) * d;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        int col = (c * 32 + lane) * VEC;
        if (col < d) {
          float t[VEC];
          ld_vec_rw<VEC>(src + col, t);
#pragma unroll
          for (int v = 0; v < VEC; ++v) acc[c][v] += t[v];
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      int col = (c * 32 + lane) * VEC;
      if (col < d) st_vec<VEC>(red + w * d + col, acc[c]);
    }
    __syncthreads();
    if (w == 0) {
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        int col = (c * 32 + lane) * VEC;
        if (col < d) {
          for (int ww = 1; ww < 8; ++ww) {
            float t[VEC];
            ld_vec_rw<VEC>(red + ww * d + col, t);
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[c][v] += t[v];
          }
        }
      }
      int n = 0;
      for (int k = lane; k < nch; k += 32) n += a.partial_occ[first + k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(kFull, n, o);
      finish_row<VEC, MAXC, UPDATE, SPEC>(a, seg, a.seg_key[seg], n, lane, acc, spec_smem, 0, rc);
    }
  }
}

template <int VEC, int MAXC, int BATCH, bool SPEC>
static void launch_seg_reduce_s(const SegArgs &a, bool update, int blocks, cudaStream_t st) {
  size_t sm1 = SPEC ? spec_smem_bytes(a.d, 8) : 0;
  if (update && VEC == 4 && MAXC == 2 && seg_bulk_ok(a)) {
    constexpr int SL = SKGE_SEG_BULK_SLOTS;
    const size_t smb = seg_bulk_smem_bytes(SL);
    cudaFuncSetAttribute(seg_reduce_bulk_kernel<SPEC, SL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb);
    int bb = kNumSMs * SKGE_SEG_BULK_CTAS;
    if (bb > blocks) bb = blocks;
    seg_reduce_bulk_kernel<SPEC, SL><<<bb, 256, smb, st>>>(a);
  } else if (update) {
    if (sm1 > 48 * 1024)
      cudaFuncSetAttribute(seg_reduce_kernel<VEC, MAXC, true, BATCH, SPEC>,
                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1);
    seg_reduce_kernel<VEC, MAXC, true, BATCH, SPEC><<<blocks, 256, sm1, st>>>(a);
  } else {
    seg_reduce_kernel<VEC, MAXC, false, BATCH, false><<<blocks, 256, 0, st>>>(a);
  }
  int cb = (a.long_chunk_cap + 7) / 8;
  if (cb > kNumSMs * 8) cb = kNumSMs * 8;
  seg_long_chunks_kernel<VEC, MAXC, BATCH><<<cb, 256, 0, st>>>(a);
  int sb = a.long_seg_cap < kNumSMs * 4 ? a.long_seg_cap : kNumSMs * 4;
  size_t smem = (size_t)8 * a.d * sizeof(float) + sm1;
  if (update) {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(seg_long_finish_kernel<VEC, MAXC, true, SPEC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)smem);
    seg_long_finish_kernel<VEC, MAXC, true, SPEC><<<sb, 256, smem, st>>>(a);
  } else {
    seg_long_finish_kernel<VEC, MAXC, false, false><<<sb, 256, smem, st>>>(a);
  }
}

// The spectral (HolE, frequency-domain) variant carries the warp FFTs; keeping it a separate
// instantiation leaves the plain update kernel with fewer registers.
template <int VEC, int MAXC, int BATCH>
static void launch_seg_reduce_b(const SegArgs &a, bool update, int blocks, cudaStream_t st) {
  if (update && a.spec_d > 0) launch_seg_reduce_s<VEC, MAXC, BATCH, true>(a, update, blocks, st);
  else launch_seg_reduce_s<VEC, MAXC, BATCH, false>(a, update, blocks, st);
}

template <int VEC, int MAXC>
static void launch_seg_reduce(const SegArgs &a, bool update, int blocks, cudaStream_t st) {
  // small minibatches (short chunks) are latency bound: more loads in flight per warp
  constexpr int BIG = MAXC <= 2 ? 4 : (MAXC == 4 ? 2 : 1);
  // (on the large minibatches 2 loads in flight, and a software-pipelined variant that fetches the next
  // segment's table entry / payloads and the parameter row ahead of the gradient rows, were both
  // measured slower than this: they cost registers, and with ~900 instructions per segment the
  // kernel lives on resident warps, not on per-warp memory parallelism)
  if (a.seg_chunk <= 32) launch_seg_reduce_b<VEC, MAXC, BIG>(a, update, blocks, st);
  else launch_seg_reduce_b<VEC, MAXC, 1>(a, update, blocks, st);
}

template <int VEC>
static int dispatch_seg_reduce(const SegArgs &a, bool update, int blocks, cudaStream_t st) {
  int chunks = (a.d + 32 * VEC - 1) / (32 * VEC);
  if (chunks <= 1) launch_seg_reduce<VEC, 1>(a, update, blocks, st);
  else if (chunks <= 2) launch_seg_reduce<VEC, 2>(a, update, blocks, st);
  else if (chunks <= 4) launch_seg_reduce<VEC, 4>(a, update, blocks, st);
  else if (chunks <= 8) launch_seg_reduce<VEC, 8>(a, update, blocks, st);
  else {
    set_error("row length %d not supported by the warp-per-row update (max %d)", a.d, 256 * VEC);
    return SKGE_EINVAL;
  }
  return 0;
}

int seg_run(const RoleMap &rm, const uint8_t *flags, int64_t P, int64_t N, int64_t M, int d,
            const float *G, int rows_per_unit, const ParamDesc pd[2], bool update, int opt, float lr,
            int32_t *counts, Arena &ar, cudaStream_t st, int spectral) {
  SegLists sl;
  int rc = seg_build(rm, flags, P, N, M, ar, st, &sl);
  if (rc) return rc;
  int64_t L = (int64_t)rm.nroles * P;
  SegArgs a;
  a.seg_chunk = seg_chunk_for(L);
  a.long_seg_cap = (int)long_seg_cap(L);
  a.long_chunk_cap = (int)long_chunk_cap(L);
  a.long_meta = ar.take<int32_t>(4);
  a.long_seg = ar.take<int32_t>((size_t)a.long_seg_cap * 3);
  a.long_work = ar.take<int32_t>((size_t)a.long_chunk_cap * 2);
  a.partials = ar.take<float>((size_t)a.long_chunk_cap * d);
  a.partial_occ = ar.take<int32_t>((size_t)a.long_chunk_cap);
  if (!ar.ok()) {
    set_error("workspace too small: need %zu bytes, have %zu", ar.off, ar.cap);
    return SKGE_EWORKSPACE;
  }
  SKGE_CUDA(cudaMemsetAsync(a.long_meta, 0, 16, st));
  a.runw = rm.runw;
  a.runw_role = rm.runw_role;
  a.coef = rm.coef;
  SKGE_REQUIRE(P * (int64_t)rows_per_unit < ((int64_t)1 << 31), "minibatch too large");
  a.spec_d = 0;
  if (spectral) {
    a.spec_d = d;
    SKGE_REQUIRE(update && spectral_len_ok(d), "spectral mode needs update mode and an even d in [32, 1024] with d / 2 = 2^a 3^b 5^c");
  }
  a.seg_start = sl.seg_start;
  a.seg_key = sl.seg_key;
  a.vals = sl.vals;
  a.meta = sl.meta;
  a.G = G;
  a.rows_per_unit = rows_per_unit;
  a.d = d;
  a.N = (int)N;
  for (int r = 0; r < 8; ++r) {
    a.grow[r] = r < rm.nroles ? rm.grow[r] : 0;
    a.gsign[r] = r < rm.nroles ? rm.gsign[r] : 0.f;
  }
  a.pd[0] = pd[0];
  a.pd[1] = pd[1];
  a.opt = opt;
  a.lr = lr;
  a.counts = counts;
  int64_t maxseg = L < N + M ? L : N + M;
  int64_t blocks = (maxseg + 7) / 8;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  if (blocks < 1) blocks = 1;
  int vec = pick_vec(d);
  // the <4, 2> spectral instantiation is the register-resident d = 256 transform: other row lengths
  // of that shape (129..255 floats, e.g. d = 200) take 64-bit accesses and the shared-memory transform
  if (a.spec_d > 0 && vec == 4 && d > 128 && d < 256) vec = 2;
  switch (vec) {
    case 4: rc = dispatch_seg_reduce<4>(a, update, (int)blocks, st); break;
    case 2: rc = dispatch_seg_reduce<2>(a, update, (int)blocks, st); break;
    default: rc = dispatch_seg_reduce<1>(a, update, (int)blocks, st); break;
  }
  if (rc) return rc;
  SKGE_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------
// standalone ParameterUpdate.__call__(g, idx) and post-hooks
// ---------------------------------------------------------------------------

template <int VEC, int MAXC>
__global__ void __launch_bounds__(256) sparse_update_warp_kernel(float *param, float *p2,
                                                                 const float *__restrict__ g,
                                                                 const int32_t *__restrict__ idx, int64_t U,
                                                                 int d, int opt, float lr, int post,
                                                                 int32_t *upd_counts) {
  const int lane = threadIdx.x & 31;
  int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t u = warp; u < U; u += nwarps) {
    int64_t row = idx ? idx[u] : u;
    float gr[MAXC][VEC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      int col = (c * 32 + lane) * VEC;
      if (col < d) {
        if (g) ld_vec<VEC>(g + u * d + col, gr[c]);
        else
#pragma unroll
          for (int v = 0; v < VEC; ++v) gr[c][v] = 0.f;
      }
    }
    row_update<VEC, MAXC>(param + row * d, p2 ? p2 + row * d : nullptr, gr, d, lane, opt, lr, post, 0.f);
    if (upd_counts && lane == 0) upd_counts[row] += 1;
  }
}

// Long rows (RESCAL's W: d*d floats): one CTA per row, two passes when a post-hook is set.
__global__ void __launch_bounds__(256) sparse_update_block_kernel(float *param, float *p2,
                                                                  const float *__restrict__ g,
                                                                  const int32_t *__restrict__ idx, int64_t U,
                                                                  int64_t rowlen, int opt, float lr, int post,
                                                                  int32_t *upd_counts) {
  __shared__ float red[40];
  for (int64_t u = blockIdx.x; u < U; u += gridDim.x) {
    int64_t row = idx ? idx[u] : u;
    float *x = param + row * rowlen;
    float *a2 = p2 ? p2 + row * rowlen : nullptr;
    const float *gu = g ? g + u * rowlen : nullptr;
    float ss = 0.f;
    for (int64_t c = threadIdx.x; c < rowlen; c += blockDim.x) {
      float xv = x[c];
      float gg = gu ? __ldg(gu + c) : 0.f;
      if (opt == SKGE_OPT_ADAGRAD) {
        float a = a2[c] + gg * gg;
        a2[c] = a;
        xv -= lr * gg * adagrad_rscale(a);
      } else {
        xv -= lr * gg;
      }
      x[c] = xv;
      ss += xv * xv;
    }
    if (post != SKGE_POST_NONE) {
      __syncthreads();
      // block_sum inline (red is static here)
      ss = warp_sum(ss);
      if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
      __syncthreads();
      if (threadIdx.x < 32) {
        float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        t = warp_sum(t);
        if (threadIdx.x == 0) red[32] = t;
      }
      __syncthreads();
      float tot = red[32];
      float scale = post == SKGE_POST_NORMALIZE ? rsqrtf(tot) : 1.0f / (tot < 1.0f ? 1.0f : tot);
      for (int64_t c = threadIdx.x; c < rowlen; c += blockDim.x) x[c] *= scale;
      __syncthreads();
    }
    if (upd_counts && threadIdx.x == 0) upd_counts[row] += 1;
  }
}

template <int VEC>
static bool launch_sparse_warp(float *param, float *p2, const float *g, const int32_t *idx, int64_t U, int d,
                               int opt, float lr, int post, int32_t *uc, cudaStream_t st) {
  int chunks = (d + 32 * VEC - 1) / (32 * VEC);
  int64_t blocks = (U + 7) / 8;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  if (chunks <= 1) sparse_update_warp_kernel<VEC, 1><<<(int)blocks, 256, 0, st>>>(param, p2, g, idx, U, d, opt, lr, post, uc);
  else if (chunks <= 2) sparse_update_warp_kernel<VEC, 2><<<(int)blocks, 256, 0, st>>>(param, p2, g, idx, U, d, opt, lr, post, uc);
  else if (chunks <= 4) sparse_update_warp_kernel<VEC, 4><<<(int)blocks, 256, 0, st>>>(param, p2, g, idx, U, d, opt, lr, post, uc);
  else if (chunks <= 8) sparse_update_warp_kernel<VEC, 8><<<(int)blocks, 256, 0, st>>>(param, p2, g, idx, U, d, opt, lr, post, uc);
  else return false;
  return true;
}

static int sparse_update_impl(float *param, float *p2, const float *g, const int32_t *idx, int64_t U,
                              int64_t rowlen, int opt, float lr, int post, int32_t *uc, cudaStream_t st) {
  bool done = false;
  if (rowlen <= 1024) {
    int d = (int)rowlen;
    switch (pick_vec(rowlen)) {
      case 4: done = launch_sparse_warp<4>(param, p2, g, idx, U, d, opt, lr, post, uc, st); break;
      case 2: done = launch_sparse_warp<2>(param, p2, g, idx, U, d, opt, lr, post, uc, st); break;
      default: done = launch_sparse_warp<1>(param, p2, g, idx, U, d, opt, lr, post, uc, st); break;
    }
  }
  if (!done) {
    int64_t blocks = U > kNumSMs * 8 ? kNumSMs * 8 : U;
    sparse_update_block_kernel<<<(int)blocks, 256, 0, st>>>(param, p2, g, idx, U, rowlen, opt, lr, post, uc);
  }
  SKGE_LAUNCH_CHECK();
  return 0;
}

}  // namespace skge

using namespace skge;

extern "C" {

int skge_sparse_update(float *param, float *p2, const float *g, const int32_t *idx, int64_t U,
                       int64_t rowlen, int opt, float lr, int post, int32_t *upd_counts,
                       skge_stream_t stream) {
  if (U == 0) return 0;
  SKGE_REQUIRE(param && g && rowlen > 0 && U > 0, "bad arguments");
  SKGE_REQUIRE(opt == SKGE_OPT_SGD || (opt == SKGE_OPT_ADAGRAD && p2), "AdaGrad needs p2");
  return sparse_update_impl(param, p2, g, idx, U, rowlen, opt, lr, post, upd_counts, as_stream(stream));
}

int skge_rows_post(float *param, const int32_t *idx, int64_t U, int64_t rowlen, int post,
                   skge_stream_t stream) {
  SKGE_REQUIRE(param && rowlen > 0 && U >= 0, "bad arguments");
  if (U == 0 || post == SKGE_POST_NONE) return 0;
  // SGD with a null gradient leaves the row unchanged and then applies the post-hook
  return sparse_update_impl(param, nullptr, nullptr, idx, U, rowlen, SKGE_OPT_SGD, 0.f, post, nullptr,
                            as_stream(stream));
}

}  // extern "C"
