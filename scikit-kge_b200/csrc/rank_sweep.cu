// Filtered ranking, CUDA-core coarse sweep (TransE L1 -- skge/run_transe.py:13-29 -- and dot-product
// models with d > 256), Blackwell version: operands are pre-packed into k-major tiles so that one
// pipeline stage is ONE contiguous bulk-TMA copy per operand, one elected lane keeps a 4-stage
// mbarrier ring full, and the eight warps run nothing but the two-instruction inner product
//   L1:  acc += |q_k - e_k|   (FADD, FADD with |.| source modifier)
//   DOT: acc += q_k * e_k     (FFMA)
// on 8 x 8 register tiles.  Rank definition: skge/base.py:950-980, 994-1017.
//
// Packed layout (skge_rank_sweep_pack): [tile of 128 rows][chunk of KC = 16 k][k][row] floats, 8 KB per
// (tile, chunk) block, zero padded in both directions.  A thread owns rows {4a..4a+3, 64+4a..64+4a+3}
// of each operand, so an operand is two 128-bit shared-memory loads per k, the sixteen lanes that
// share a query row read sixteen consecutive 16-byte pieces (no bank conflict), and only 16 operand
// registers are live next to the 64 accumulators: no spills at two CTAs per SM.
#include "common.cuh"
#include "umma.cuh"

namespace skge {
namespace sw {

using namespace ptx;

static constexpr int TILE = 128;                  // rows per operand tile
static constexpr int KC = 16;                     // k per stage
static constexpr int BLOCK_FLOATS = TILE * KC;    // one (tile, chunk) block
static constexpr int BLOCK_BYTES = BLOCK_FLOATS * 4;
static constexpr int NSTAGE = 4;
static constexpr int COMPUTE_WARPS = 8;
static constexpr int THREADS = 32 * COMPUTE_WARPS;

struct __align__(16) Smem {
  float q[NSTAGE][BLOCK_FLOATS];
  float e[NSTAGE][BLOCK_FLOATS];
  float thi[TILE], tlo[TILE];
  uint64_t full[NSTAGE], empty[NSTAGE];
};

// src [rows][d] row-major -> packed blocks.  One CTA per (tile, chunk), 128 threads: thread r reads
// KC consecutive floats of row r (64 contiguous bytes) and writes them transposed (coalesced over r).
__global__ void __launch_bounds__(TILE) sweep_pack_kernel(const float *__restrict__ src, int64_t rows, int d, int nch,
                                                          float *__restrict__ out) {
  const int64_t tile = blockIdx.x / nch;
  const int c = blockIdx.x - tile * nch;
  const int r = threadIdx.x;
  const int64_t row = tile * TILE + r;
  float v[KC];
#pragma unroll
  for (int k = 0; k < KC; ++k) {
    const int kk = c * KC + k;
    v[k] = (row < rows && kk < d) ? __ldg(src + row * d + kk) : 0.f;
  }
  float *o = out + ((int64_t)blockIdx.x) * BLOCK_FLOATS;
#pragma unroll
  for (int k = 0; k < KC; ++k) o[k * TILE + r] = v[k];
}

template <int OP>
__global__ void __launch_bounds__(THREADS, 2) rank_sweep_tma_kernel(const float *__restrict__ Epk, int64_t n_shard,
                                                                    int64_t shard_base, int d, int nch,
                                                                    const float *__restrict__ Qpk,
                                                                    const double *__restrict__ tscore,
                                                                    const float *__restrict__ eps, int64_t Q,
                                                                    int32_t *__restrict__ cnt_gt,
                                                                    int32_t *__restrict__ cand_q,
                                                                    int32_t *__restrict__ cand_e, int64_t cand_cap,
                                                                    unsigned long long *__restrict__ cand_count,
                                                                    int etiles_per_cta) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t q0 = (int64_t)blockIdx.x * TILE;
  const int64_t ntiles_e = (n_shard + TILE - 1) / TILE;
  const int64_t et_beg = (int64_t)blockIdx.y * etiles_per_cta;
  const int64_t et_end = min(ntiles_e, et_beg + etiles_per_cta);

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], COMPUTE_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (threadIdx.x < TILE) {
    const int64_t q = q0 + threadIdx.x;
    float hi = INFINITY, lo = INFINITY;   // rows beyond Q: never counted, never in the band
    if (q < Q) {
      const double t = tscore[q], e = (double)eps[q];
      hi = __double2float_ru(t + e);
      lo = __double2float_rd(t - e);
    }
    sm.thi[threadIdx.x] = hi;
    sm.tlo[threadIdx.x] = lo;
  }
  __syncthreads();

  // Chunk i of this CTA's stream = (entity tile et_beg + i / nch, k chunk i % nch); it lives in stage
  // i % NSTAGE.  Warp 0 issues chunk i + NSTAGE - 1 before it computes chunk i: that stage held chunk
  // i - 1, so the wait on its "empty" barrier couples the warps with one chunk of slack.
  const int64_t nchunks = (et_end - et_beg) * nch;
  const float *qsrc = Qpk + (int64_t)blockIdx.x * nch * BLOCK_FLOATS;
  const float *esrc = Epk + et_beg * nch * BLOCK_FLOATS;   // the CTA's entity chunks are contiguous
  auto issue = [&](int64_t i) {
    const uint32_t s = (uint32_t)(i % NSTAGE), ph = (uint32_t)((i / NSTAGE) & 1);
    mbar_wait(&sm.empty[s], ph ^ 1);
    if (elect_one()) {
      mbar_expect_tx(&sm.full[s], 2 * BLOCK_BYTES);
      bulk_g2s(sm.q[s], qsrc + (i % nch) * BLOCK_FLOATS, BLOCK_BYTES, &sm.full[s]);
      bulk_g2s(sm.e[s], esrc + i * BLOCK_FLOATS, BLOCK_BYTES, &sm.full[s]);
    }
    __syncwarp();
  };
  if (warp == 0)
    for (int64_t i = 0; i < NSTAGE - 1 && i < nchunks; ++i) issue(i);

  const int klast = ((d - (nch - 1) * KC) + 3) & ~3;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // entity rows 4 tx.., query rows 4 ty..
  uint32_t stage = 0, phase = 0;
  int cnt[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) cnt[i] = 0;

  int64_t chunk = 0;
  for (int64_t et = et_beg; et < et_end; ++et) {
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int c = 0; c < nch; ++c, ++chunk) {
      if (warp == 0 && chunk + NSTAGE - 1 < nchunks) issue(chunk + NSTAGE - 1);
      mbar_wait(&sm.full[stage], phase);
      const float *pq = sm.q[stage] + 4 * ty, *pe = sm.e[stage] + 4 * tx;
      // 4 k per loop body: 512 math instructions = 8 KB of code.  The warps of a CTA are not in lockstep
      // here (no CTA barrier in the loop), so a fully unrolled stage (33 KB) thrashes the instruction
      // cache: ncu showed 60 % of the stall samples as "no instruction".
      const int kend = c == nch - 1 ? klast : KC;   // the zero padding of the last chunk is skipped (4 k granularity)
#pragma unroll 4
      for (int k = 0; k < kend; ++k) {
        const float4 qa = *reinterpret_cast<const float4 *>(pq + k * TILE);
        const float4 qb = *reinterpret_cast<const float4 *>(pq + k * TILE + 64);
        const float4 ea = *reinterpret_cast<const float4 *>(pe + k * TILE);
        const float4 eb = *reinterpret_cast<const float4 *>(pe + k * TILE + 64);
        const float qv[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
        const float ev[8] = {ea.x, ea.y, ea.z, ea.w, eb.x, eb.y, eb.z, eb.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (OP == SKGE_RANK_L1) acc[i][j] += fabsf(qv[i] - ev[j]);
            else acc[i][j] = fmaf(qv[i], ev[j], acc[i][j]);
          }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.empty[stage]);
      if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
    }
    // epilogue: compare against the per-query thresholds; a score is never stored.  Two predicate
    // bits per accumulator (bit j of `gt`: counted, of `ge`: at or above the band's lower edge), one
    // popc per query row; the rare band elements (ge & ~gt) are pushed from a loop over the set bits,
    // so the code stays a few KB: with the candidate push inlined per element the epilogue was
    // 46 KB of straight-line code that every warp walked once per tile and that evicted the 8 KB
    // loop body from the instruction cache (ncu at config 1: 38 % of the stall samples there).
    const int nvalid = (int)min((int64_t)TILE, n_shard - et * TILE);
    const int ebase = (int)(shard_base + et * TILE);
    // valid entity rows of this thread: 4 tx + j (j < 4) and 64 + 4 tx + (j - 4)
    const unsigned vm = ((1u << min(4, max(0, nvalid - 4 * tx))) - 1u) |
                        (((1u << min(4, max(0, nvalid - 64 - 4 * tx))) - 1u) << 4);
    unsigned long long bands = 0;   // bit 8 i + j: accumulator (i, j) is inside its query's band
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int qr = (i < 4 ? 0 : 60) + 4 * ty + i;          // rows 4 ty + i, 64 + 4 ty + (i - 4)
      // L1: score = -acc, so score > thi <=> acc < -thi and score >= tlo <=> acc <= -tlo (exact)
      const float thi = OP == SKGE_RANK_L1 ? -sm.thi[qr] : sm.thi[qr];
      const float tlo = OP == SKGE_RANK_L1 ? -sm.tlo[qr] : sm.tlo[qr];
      unsigned gt = 0, ge = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (OP == SKGE_RANK_L1) {
          gt |= acc[i][j] < thi ? (1u << j) : 0u;
          ge |= acc[i][j] <= tlo ? (1u << j) : 0u;
        } else {
          gt |= acc[i][j] > thi ? (1u << j) : 0u;
          ge |= acc[i][j] >= tlo ? (1u << j) : 0u;
        }
      }
      gt &= vm;
      cnt[i] += __popc(gt);
      bands |= (unsigned long long)(ge & ~gt & vm) << (8 * i);
    }
    while (bands) {   // rare: a handful of pairs per query over the whole table
      const int b = __ffsll((long long)bands) - 1;
      bands &= bands - 1;
      const int i = b >> 3, j = b & 7;
      const unsigned long long slot = atomicAdd(cand_count, 1ull);
      if ((int64_t)slot < cand_cap) {
        cand_q[slot] = (int32_t)(q0 + (i < 4 ? 0 : 60) + 4 * ty + i);
        cand_e[slot] = ebase + (j < 4 ? 0 : 60) + 4 * tx + j;
      }
    }
  }
  // the 16 threads sharing ty sit in one half-warp: reduce, then one atomic per query
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int c = cnt[i];
    c += __shfl_xor_sync(kFull, c, 8);
    c += __shfl_xor_sync(kFull, c, 4);
    c += __shfl_xor_sync(kFull, c, 2);
    c += __shfl_xor_sync(kFull, c, 1);
    const int64_t q = q0 + (i < 4 ? 0 : 60) + 4 * ty + i;
    if (tx == 0 && q < Q && c) atomicAdd(cnt_gt + q, c);
  }
}

}  // namespace sw
}  // namespace skge

using namespace skge;
using namespace skge::sw;

extern "C" {

int64_t skge_rank_sweep_packed_floats(int64_t rows, int d) {
  if (rows < 0 || d <= 0) return -1;
  return ((rows + TILE - 1) / TILE) * (int64_t)((d + KC - 1) / KC) * BLOCK_FLOATS;
}

int skge_rank_sweep_pack(const float *src, int64_t rows, int d, float *out, skge_stream_t stream) {
  SKGE_REQUIRE(src && out && rows >= 0 && d > 0, "bad arguments");
  if (rows == 0) return 0;
  const int nch = (d + KC - 1) / KC;
  const int64_t blocks = ((rows + TILE - 1) / TILE) * nch;
  SKGE_REQUIRE(blocks < (1ll << 31), "table too large for one pack launch");
  sweep_pack_kernel<<<(unsigned)blocks, TILE, 0, as_stream(stream)>>>(src, rows, d, nch, out);
  SKGE_LAUNCH_CHECK();
  return 0;
}

int skge_rank_sweep_tiles(int op, const float *Epk, int64_t n_shard, int64_t shard_base, int d, const float *Qpk,
                          const double *tscore, const float *eps, int64_t Q, int32_t *cnt_gt, int32_t *cand_q,
                          int32_t *cand_e, int64_t cand_cap, unsigned long long *cand_count, skge_stream_t stream) {
  SKGE_REQUIRE(Epk && Qpk && tscore && eps && cnt_gt && cand_q && cand_e && cand_count, "null argument");
  SKGE_REQUIRE((op == SKGE_RANK_L1 || op == SKGE_RANK_DOT) && d > 0 && n_shard >= 0 && Q >= 0, "bad sizes");
  if (Q == 0 || n_shard == 0) return 0;
  const int nch = (d + KC - 1) / KC;
  const int64_t qtiles = (Q + TILE - 1) / TILE, etiles = (n_shard + TILE - 1) / TILE;
  // Split the entity range so that the CTAs fill whole waves of the GPU (2 resident CTAs per SM):
  // the smallest split whose wave efficiency is >= 97 %, else the best one.
  const int64_t slots = 2 * kNumSMs;
  int64_t ysplit = 1;
  int per = (int)etiles;
  double best = -1.0;
  for (int64_t ys = 1; ys <= etiles && ys <= 64; ++ys) {
    const int64_t p = (etiles + ys - 1) / ys, yeff = (etiles + p - 1) / p;
    if (yeff != ys) continue;  // same partition as a smaller split
    const int64_t total = qtiles * ys, waves = (total + slots - 1) / slots;
    const double eff = (double)(qtiles * etiles) / ((double)waves * slots * p);
    if (eff > best + 1e-9) { best = eff; ysplit = ys; per = (int)p; }
    if (eff >= 0.97) break;
  }
  SKGE_REQUIRE(qtiles < (1ll << 31), "too many queries for one launch");
  dim3 grid((unsigned)qtiles, (unsigned)ysplit);
  cudaStream_t st = as_stream(stream);
  const size_t smem = sizeof(Smem);
  if (op == SKGE_RANK_L1) {
    SKGE_CUDA(cudaFuncSetAttribute(rank_sweep_tma_kernel<SKGE_RANK_L1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    rank_sweep_tma_kernel<SKGE_RANK_L1><<<grid, THREADS, smem, st>>>(Epk, n_shard, shard_base, d, nch, Qpk, tscore,
                                                                    eps, Q, cnt_gt, cand_q, cand_e, cand_cap, cand_count,
                                                                    per);
  } else {
    SKGE_CUDA(cudaFuncSetAttribute(rank_sweep_tma_kernel<SKGE_RANK_DOT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    rank_sweep_tma_kernel<SKGE_RANK_DOT><<<grid, THREADS, smem, st>>>(Epk, n_shard, shard_base, d, nch, Qpk, tscore,
                                                                     eps, Q, cnt_gt, cand_q, cand_e, cand_cap,
                                                                     cand_count, per);
  }
  SKGE_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
