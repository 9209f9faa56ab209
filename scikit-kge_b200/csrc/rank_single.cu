// HolE / RESCAL filtered ranking, large sweeps: ONE fp16 product on the tensor cores + an int8
// refinement of BOTH cross terms in the epilogue  (skge/run_hole.py:15-19 as one GEMM whose result
// is never stored; rank definition skge/base.py:950-980, 994-1017).
//
//   score[q][e] - tmid[q] = q_hi.e_hi - tmid[q]        (tcgen05.mma, fp32 accumulators in TMEM)
//                         + q_hi.e_lo + q_lo.e_hi      (only where it can change the outcome)
//
// Every operand is split x * scale = hi + lo with hi, lo in fp16.  The two missing products are
// bounded by ||q|| max||e_lo|| + ||q_lo|| max(||e_hi|| + ||e_lo||) over the 128 packed rows of an
// entity tile (rows are packed by decreasing norm, so a tile's rows are alike).  What the epilogue
// does per accumulator element is therefore only
//   * a sign test (is the coarse score above the middle of the query's undecided band?), folded
//     into one LEA.HI per element because the accumulator is PRE-LOADED with -tmid[q]
//     (tcgen05.st by the epilogue warp that drained the stage, so every MMA accumulates), and
//   * |acc| <= h[q][tile] for the wide band, as a running 3-input minimum (FMNMX3 with |.|): half an
//     instruction per element; the compare happens once per 32 elements.
// The ~0.5 % of pairs inside the wide band get both cross terms from 8-bit copies of all four
// vectors (dp4a, exact integer arithmetic, quantisation error bounded per pair) and are then
// tested against the tight band; what stays undecided goes to the candidate list that
// skge_rank_rescore settles in fp64, so no result depends on the low-precision arithmetic.
//
// Kernel shape: as csrc/rank_refine.cu (persistent CTA or cta_group::2 pair per SM, 20 warps:
// bulk-TMA producer, MMA issuer, relay, 16 epilogue warps), UMMA N = 256, 8 KB stages.
#include <cuda_fp16.h>

#include "common.cuh"
#include "umma.cuh"

namespace skge {
namespace rs {

using namespace ptx;

#ifndef SKGE_EPI_SLEEP
#define SKGE_EPI_SLEEP 128u
#endif
static constexpr int QT = 128;             // query rows per CTA (UMMA M per CTA)
static constexpr int ET = 256;             // entity rows per MMA (UMMA N)
static constexpr int BLOCK_BYTES = 16384;  // one (128-row tile, 64-k chunk) fp16 block
static constexpr int STAGE_BYTES = 8192;   // B stage: 256 rows x 16 k (CG = 1) or 128 rows x 32 k per CTA (CG = 2)
static constexpr int MAX_KCH = 4;          // d <= 256
static constexpr int MAX_NB = 10;
static constexpr int LIST_CAP = 48;        // wide-band pairs per warp and tile half
static constexpr int THREADS = 640;
static constexpr int QM = 16;              // floats of per-query constants

struct __align__(8) Ctrl {
  uint64_t a_full, a_peer, a_empty;
  uint64_t b_full[MAX_NB], b_peer[MAX_NB], b_empty[MAX_NB];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_base, pad;
};
struct Lists {
  uint2 ent[16][LIST_CAP];   // (accumulator bits, lane << 8 | column)
  int count[16];
};

struct SingleArgs {
  const uint8_t *Ehi;        // fp16 blocks, an even number of 128-row tiles (zero padded)
  const int8_t *E8;          // [packed rows][2][kb]: int8 lo row, int8 hi row
  const float4 *e_meta;      // [packed rows] (scale_lo, l1_lo, scale_hi, l1_hi)
  const float2 *tile_w;      // [2 * etiles] per 128-row tile: (max ||e_lo||, max (||e_hi|| + ||e_lo||)), rounded up
  const int32_t *perm;       // nullable: packed row -> shard-local entity id
  const uint8_t *Qhi;        // fp16 blocks [qtiles][kch]
  const int8_t *Q8h, *Q8l;   // [qtiles][128][kb], swizzled
  const float *qmeta;        // [qtiles * 128][QM]: tmid, htight, ||q||, ||q_lo||, sqh, sql, qA1, qB1, qA2, qB2, 0...
  int64_t n_shard, shard_base, Q;
  int kch, nb, qtiles, qunits, etiles, nslices, tiles_per_slice;
  int32_t *cnt_gt, *cand_q, *cand_e;
  int64_t cand_cap;
  unsigned long long *cand_count;
};

struct Item { int qunit, et_beg, et_end; };
__device__ __forceinline__ Item get_item(const SingleArgs &a, int item) {
  Item it;
  const int slice = item / a.qunits;
  it.qunit = item - slice * a.qunits;
  it.et_beg = slice * a.tiles_per_slice;
  it.et_end = min(a.etiles, it.et_beg + a.tiles_per_slice);
  return it;
}

__device__ __forceinline__ void push_global(const SingleArgs &a, int q, int e) {
  unsigned long long slot = atomicAdd(a.cand_count, 1ull);
  if ((int64_t)slot < a.cand_cap) {
    a.cand_q[slot] = q;
    a.cand_e[slot] = e;
  }
}
__device__ __forceinline__ int64_t entity_of(const SingleArgs &a, int64_t r) {
  return a.perm ? (int64_t)__ldg(a.perm + r) : r;
}

// r[j] for a run-time j without spilling the array: a 5-level multiplexer
__device__ __forceinline__ uint32_t pick32(const uint32_t (&r)[32], int j) {
  uint32_t a[16], b[8], c[4];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (j & 1) ? r[2 * i + 1] : r[2 * i];
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = (j & 2) ? a[2 * i + 1] : a[2 * i];
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i] = (j & 4) ? b[2 * i + 1] : b[2 * i];
  const uint32_t d0 = (j & 8) ? c[1] : c[0], d1 = (j & 8) ? c[3] : c[2];
  return (j & 16) ? d1 : d0;
}

// eight TMEM columns <- eight registers per lane.  STTM takes a block of consecutive registers, so the
// caller keeps eight copies of the fill value alive (made opaque with copy8, otherwise ptxas merges
// them and re-materialises the copies with eight MOVs in front of every store).
__device__ __forceinline__ void tmem_fill8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void copy8(uint32_t (&v)[8], uint32_t x) {
#pragma unroll
  for (int i = 0; i < 8; ++i) asm volatile("mov.b32 %0, %1;" : "=r"(v[i]) : "r"(x));
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int CG, int KCH>
__global__ void __launch_bounds__(THREADS, 1) rank_single_kernel(const SingleArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  constexpr int kch = KCH, kb = KCH * 64;   // compile-time row length: the refinement's addressing folds into immediates
  uint8_t *sA_hi = smem_raw;
  uint8_t *sQ8h = sA_hi + kch * BLOCK_BYTES;
  uint8_t *sQ8l = sQ8h + QT * kb;
  float *sQm = reinterpret_cast<float *>(sQ8l + QT * kb);
  uint8_t *sB = reinterpret_cast<uint8_t *>(sQm + QT * QM);
  Ctrl *ctrl = reinterpret_cast<Ctrl *>(sB + a.nb * STAGE_BYTES);
  Lists *wl = reinterpret_cast<Lists *>(ctrl + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const int unit = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int nunits = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int nitems = a.qunits * a.nslices;
  const int nb = a.nb;
  // B stages per 256-row entity tile: 16 k each (CG = 1) or 32 k each (CG = 2)
  const int nks = CG == 2 ? kch * 2 : kch * 4;

  if (threadIdx.x == 0) {
    mbar_init(&ctrl->a_full, 1);
    mbar_init(&ctrl->a_peer, 1);
    mbar_init(&ctrl->a_empty, 1 + 16);   // MMA commit + the 16 epilogue warps (they read the int8 rows)
    for (int s = 0; s < nb; ++s) {
      mbar_init(&ctrl->b_full[s], 1);
      mbar_init(&ctrl->b_peer[s], 1);
      mbar_init(&ctrl->b_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&ctrl->acc_full[s], 1);
      mbar_init(&ctrl->acc_empty[s], 8 * CG);   // two quads per accumulator stage, in each CTA
    }
    for (int w = 0; w < 16; ++w) wl->count[w] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<CG>(&ctrl->tmem_base, 512u);
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();   // the peer's barriers exist before anything arrives on them
  tc_fence_after();
  const uint32_t tmem = ctrl->tmem_base;

  if (warp == 0) {
    // ===================== producer: bulk copies global -> shared =====================
    uint32_t bstage = 0, bphase = 0, aphase = 0;
    const uint32_t a_bytes = (uint32_t)(kch * BLOCK_BYTES + 2 * QT * kb + QT * QM * 4);
    for (int item = unit; item < nitems; item += nunits) {
      const Item it = get_item(a, item);
      const int qt = min(a.qtiles - 1, CG == 2 ? 2 * it.qunit + (int)rank : it.qunit);
      mbar_wait(&ctrl->a_empty, aphase ^ 1);  // previous item's MMAs retired, epilogue done with the int8 rows
      if (elect_one()) {
        mbar_expect_tx(&ctrl->a_full, a_bytes);
        const uint8_t *qh = a.Qhi + (int64_t)qt * kch * BLOCK_BYTES;
        for (int c = 0; c < kch; ++c)
          bulk_g2s(sA_hi + c * BLOCK_BYTES, qh + (int64_t)c * BLOCK_BYTES, BLOCK_BYTES, &ctrl->a_full);
        bulk_g2s(sQ8h, a.Q8h + (int64_t)qt * QT * kb, (uint32_t)(QT * kb), &ctrl->a_full);
        bulk_g2s(sQ8l, a.Q8l + (int64_t)qt * QT * kb, (uint32_t)(QT * kb), &ctrl->a_full);
        bulk_g2s(sQm, a.qmeta + (int64_t)qt * QT * QM, (uint32_t)(QT * QM * 4), &ctrl->a_full);
      }
      __syncwarp();
      aphase ^= 1;
      for (int et = it.et_beg; et < it.et_end; ++et) {
        // my part of the entity tile: 128-row tile 2 et + rank (CG = 2) or both 128-row tiles (CG = 1)
        const uint8_t *tile0 = a.Ehi + (int64_t)(2 * et + (CG == 2 ? (int)rank : 0)) * kch * BLOCK_BYTES;
        for (int ks = 0; ks < nks; ++ks) {
          mbar_wait(&ctrl->b_empty[bstage], bphase ^ 1);
          uint8_t *dst = sB + bstage * STAGE_BYTES;
          if (elect_one()) {
            mbar_expect_tx(&ctrl->b_full[bstage], STAGE_BYTES);
            if (CG == 2) {
              // k range [32 ks, 32 ks + 32): 8 KB, contiguous in the block
              bulk_g2s(dst, tile0 + (int64_t)ks * STAGE_BYTES, STAGE_BYTES, &ctrl->b_full[bstage]);
            } else {
              // 256 rows x 16 k as [kcore 2][rowgroup 32][8][16 B]: four 2 KB pieces of two 128-row tiles
              const uint8_t *src = tile0 + (int64_t)(ks >> 2) * BLOCK_BYTES + (ks & 3) * 4096;
              const int64_t t1 = (int64_t)kch * BLOCK_BYTES;
              bulk_g2s(dst, src, 2048, &ctrl->b_full[bstage]);
              bulk_g2s(dst + 2048, src + t1, 2048, &ctrl->b_full[bstage]);
              bulk_g2s(dst + 4096, src + 2048, 2048, &ctrl->b_full[bstage]);
              bulk_g2s(dst + 6144, src + t1 + 2048, 2048, &ctrl->b_full[bstage]);
            }
          }
          __syncwarp();
          if (++bstage == (uint32_t)nb) { bstage = 0; bphase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA) =====================
    if (leader) {
      constexpr uint32_t IDESC = idesc_f16(QT * CG, ET);
      constexpr uint32_t B_LBO = CG == 2 ? 2048u : 4096u;
      const uint64_t ahi0 = make_desc(smem_u32(sA_hi), 2048u, 128u);
      const uint64_t b00 = make_desc(smem_u32(sB), B_LBO, 128u);
      uint32_t bstage = 0, bphase = 0, aphase = 0, accs = 0, accphase = 0;
      for (int item = unit; item < nitems; item += nunits) {
        const Item it = get_item(a, item);
        mbar_wait(&ctrl->a_full, aphase);
        if (CG == 2) mbar_wait(&ctrl->a_peer, aphase);
        aphase ^= 1;
        for (int et = it.et_beg; et < it.et_end; ++et) {
          // phase k of acc_empty completes when the epilogue has pre-loaded the stage for its k-th use
          mbar_wait(&ctrl->acc_empty[accs], accphase);
          tc_fence_after();
          const uint32_t d_tmem = tmem + accs * ET;
          for (int ks = 0; ks < nks; ++ks) {
            mbar_wait(&ctrl->b_full[bstage], bphase);
            if (CG == 2) mbar_wait(&ctrl->b_peer[bstage], bphase);
            tc_fence_after();
            const uint64_t bd = b00 + (uint64_t)((bstage * STAGE_BYTES) >> 4);
            if (elect_one()) {
#pragma unroll
              for (int j = 0; j < (CG == 2 ? 2 : 1); ++j) {
                const int k16 = CG == 2 ? 2 * ks + j : ks;   // 16-k step within the row
                const uint32_t aoff = ((uint32_t)(k16 >> 2) * BLOCK_BYTES + (uint32_t)(k16 & 3) * 4096u) >> 4;
                umma_f16<CG>(d_tmem, ahi0 + aoff, bd + (uint64_t)(j * 256), IDESC, 1u);   // always accumulates
              }
              tc_commit<CG>(&ctrl->b_empty[bstage]);  // frees the stage (in both CTAs) once these MMAs have read it
              if (ks == nks - 1) tc_commit<CG>(&ctrl->acc_full[accs]);
            }
            __syncwarp();
            if (++bstage == (uint32_t)nb) { bstage = 0; bphase ^= 1; }
          }
          if (++accs == 2) { accs = 0; accphase ^= 1; }
        }
        if (elect_one()) tc_commit<CG>(&ctrl->a_empty);
        __syncwarp();
      }
    }
  } else if (warp == 2) {
    // ===================== relay (peer CTA of a pair): my copies have landed -> tell the leader =====
    if (CG == 2 && !leader) {
      uint32_t bstage = 0, bphase = 0, aphase = 0;
      for (int item = unit; item < nitems; item += nunits) {
        const Item it = get_item(a, item);
        mbar_wait(&ctrl->a_full, aphase);
        if (elect_one()) mbar_arrive_remote(&ctrl->a_peer, 0);
        __syncwarp();
        aphase ^= 1;
        for (int et = it.et_beg; et < it.et_end; ++et) {
          for (int ks = 0; ks < nks; ++ks) {
            mbar_wait(&ctrl->b_full[bstage], bphase);
            if (elect_one()) mbar_arrive_remote(&ctrl->b_peer[bstage], 0);
            __syncwarp();
            if (++bstage == (uint32_t)nb) { bstage = 0; bphase ^= 1; }
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int w16 = warp - 4, quarter = warp & 3, quad = w16 >> 2;
    const uint32_t st = (uint32_t)(quad >> 1);      // accumulator stage this quad serves
    const int colhalf = quad & 1;                    // which 128 of the stage's 256 columns
    const int row = quarter * 32 + lane;             // query row inside the tile
    const int slot = lane & 7, part = lane >> 3;     // refinement: four lanes per pair
    constexpr bool swz = (KCH & 1) == 0;
    // bank group of the swizzled shared-memory reads: distinct over the 8 slots of a quarter-warp
    const int pi[2] = {(slot + part) & 7, (slot + part + 4) & 7};
    constexpr uint32_t EPI_SLEEP = SKGE_EPI_SLEEP;   // ns between polls of a waiting epilogue warp
    const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + st * ET + colhalf * 128;

    // Look-ahead over this CTA's tile sequence: `la` is the next tile this warp's accumulator stage
    // will hold; the stage is pre-loaded with -tmid of THAT tile's query row when it is handed back.
    int la_item = unit, la_et = 0, la_end = 0;
    bool la_valid = false;
    uint32_t la_ntm[8];                              // eight copies of the bits of -tmid[row] for the look-ahead tile's item
    copy8(la_ntm, 0u);
    auto la_load = [&]() {
      la_valid = la_item < nitems;
      if (la_valid) {
        const Item it = get_item(a, la_item);
        la_et = it.et_beg;
        la_end = it.et_end;
        const int qt = CG == 2 ? 2 * it.qunit + (int)rank : it.qunit;
        const int64_t q = (int64_t)qt * QT + row;
        copy8(la_ntm, q < a.Q ? __float_as_uint(-__ldg(a.qmeta + q * QM)) : 0u);
      }
    };
    auto la_step = [&]() {
      if (la_valid && ++la_et == la_end) {
        la_item += nunits;
        la_load();
      }
    };
    auto fill_stage = [&]() {   // 128 columns of my lanes <- -tmid, then hand the stage to the MMA warp
#pragma unroll
      for (int c = 0; c < 16; ++c) tmem_fill8(taddr + 8 * c, la_ntm);
    };
    auto release_stage = [&]() {
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2 && !leader) mbar_arrive_remote(&ctrl->acc_empty[st], 0);
        else mbar_arrive(&ctrl->acc_empty[st]);
      }
    };
    la_load();
    if (st == 1) la_step();
    if (la_valid) {
      fill_stage();
      release_stage();
    }

    uint32_t tseq = 0, aphase = 0;
    for (int item = unit; item < nitems; item += nunits) {
      const Item it = get_item(a, item);
      const int qt = CG == 2 ? 2 * it.qunit + (int)rank : it.qunit;
      const int64_t q = (int64_t)qt * QT + row;
      float htight = -1.f, qn = 0.f, qlon = 0.f;    // rows beyond Q: empty band, count ignored
      if (q < a.Q) {
        const float4 m0 = __ldg(reinterpret_cast<const float4 *>(a.qmeta + q * QM));
        htight = m0.y; qn = m0.z; qlon = m0.w;
      }
      mbar_wait_sleep<EPI_SLEEP>(&ctrl->a_full, aphase);   // int8 query rows and constants of this item are in shared memory
      aphase ^= 1;
      int cnt = 0;
      for (int et = it.et_beg; et < it.et_end; ++et) {
        const uint32_t my = tseq++;
        if ((my & 1u) != st) continue;                // the other quads' accumulator stage
        la_step();                                    // la was this tile: move it to my next one
        la_step();
        const int64_t e0 = (int64_t)et * ET + colhalf * 128;
        const int nvalid = (int)max((int64_t)0, min((int64_t)128, a.n_shard - e0));
        // this tile's wide band: the two missing products are at most ||q|| max||e_lo|| + ||q_lo|| max(||e_hi|| + ||e_lo||)
        const float2 tw = __ldg(a.tile_w + 2 * et + colhalf);
        const float h = __fmaf_ru(qlon, tw.y, __fmaf_ru(qn, tw.x, htight));
        mbar_wait_sleep<EPI_SLEEP>(&ctrl->acc_full[st], (my >> 1) & 1u);
        tc_fence_after();
        int nlist = 0;
        for (int c2 = 0; c2 < 4; ++c2) {
          uint32_t r[32];
          tmem_ld32(taddr + 32 * c2, r);
          tmem_ld_wait();
          if (la_valid) {   // these 32 columns are in registers: pre-load them for the stage's next tile
#pragma unroll
            for (int c = 0; c < 4; ++c) tmem_fill8(taddr + 32 * c2 + 8 * c, la_ntm);
            if (c2 == 3) release_stage();
          }
          const int left = nvalid - c2 * 32;
          uint32_t band = 0u;
          if (left >= 32) {
            // fast path: one LEA.HI (sign count) and half an FMNMX3 (running |.| minimum) per element
            uint32_t nneg = 0u;
            float mg[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float m = fabsf(__uint_as_float(r[8 * g]));
#pragma unroll
              for (int j = 1; j < 8; j += 2)
                m = fminf(m, fminf(fabsf(__uint_as_float(r[8 * g + j])),
                                   fabsf(__uint_as_float(r[8 * g + (j + 1 < 8 ? j + 1 : j)]))));
              mg[g] = m;
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) nneg += r[j] >> 31;
            cnt += 32 - (int)nneg;
            if (fminf(fminf(mg[0], mg[1]), fminf(mg[2], mg[3])) <= h) {
#pragma unroll
              for (int g = 0; g < 4; ++g)
                if (mg[g] <= h) {
#pragma unroll
                  for (int j = 0; j < 8; ++j)
                    if (fabsf(__uint_as_float(r[8 * g + j])) <= h) band |= 1u << (8 * g + j);
                }
            }
          } else if (left > 0) {
            // ragged end of the shard: only the first `left` columns are entities
            uint32_t mpos = 0u;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if ((r[j] >> 31) == 0u) mpos |= 1u << j;
              if (fabsf(__uint_as_float(r[j])) <= h) band |= 1u << j;
            }
            const uint32_t vm = 0xFFFFFFFFu >> (32 - left);
            cnt += __popc(mpos & vm);
            band &= vm;
          }
          // Wide-band elements -> the warp's list.  Slots come from a warp prefix sum of the per-lane
          // counts (no shared-memory atomics in the dependent chain); nlist is warp-uniform.
          const uint32_t hit = __ballot_sync(kFull, band != 0u);
          if (hit) {
            const int mine_n = __popc(band);
            int idx;
            if (__ballot_sync(kFull, mine_n > 1) == 0u) {
              // the usual case, one element per lane: slots straight from the ballot
              idx = nlist + __popc(hit & ((1u << lane) - 1u));
              nlist += __popc(hit);
            } else {
              int pre_n = mine_n;
#pragma unroll
              for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(kFull, pre_n, o);
                if (lane >= o) pre_n += t;
              }
              idx = nlist + pre_n - mine_n;
              nlist += __shfl_sync(kFull, pre_n, 31);
            }
            while (band) {   // inside the WIDE band: needs the two missing products
              const int j = __ffs(band) - 1;
              band &= band - 1;
              const uint32_t bits = pick32(r, j);
              if (idx < LIST_CAP) {
                wl->ent[w16][idx] = make_uint2(bits, (uint32_t)((lane << 8) | (c2 * 32 + j)));
              } else {
                // list full (rare): the fp64 pass settles this pair; it was counted if its sign bit is clear
                if ((bits >> 31) == 0u) --cnt;
                push_global(a, (int)q, (int)(a.shard_base + entity_of(a, e0 + c2 * 32 + j)));
              }
              ++idx;
            }
          }
        }
        const int n = min(nlist, LIST_CAP);
        if (n) {
          __syncwarp();   // the list entries written above are visible to the whole warp
          for (int b = 0; b < n; b += 8) {
            const bool mine = b + slot < n;
            const uint2 en = wl->ent[w16][mine ? b + slot : 0];
            const int L = (int)(en.y >> 8), col = (int)(en.y & 255u);
            const int qr = quarter * 32 + L;
            const int64_t erow = e0 + col;
            const int8_t *ebase = a.E8 + erow * (2 * kb);
            const uint8_t *qhb = sQ8h + qr * kb, *qlb = sQ8l + qr * kb;
            const int x = swz ? (qr & 7) : 0;
            // every load of the pair is issued before the first use: the rows' chunks and the row's constants
            int4 wlo[KCH], whi[KCH];
#pragma unroll
            for (int t = 0; t < KCH; ++t) {
              const int c = swz ? (8 * (t >> 1) + (pi[t & 1] ^ x)) : part + 4 * t;
              wlo[t] = __ldg(reinterpret_cast<const int4 *>(ebase + c * 16));
              whi[t] = __ldg(reinterpret_cast<const int4 *>(ebase + kb + c * 16));
            }
            const float4 em = __ldg(a.e_meta + erow);                                    // selo, l1lo, sehi, l1hi
            int acc1 = 0, acc2 = 0;   // q_hi8 . e_lo8   and   q_lo8 . e_hi8
#pragma unroll
            for (int t = 0; t < KCH; ++t) {
              const int phi = swz ? 8 * (t >> 1) + pi[t & 1] : part + 4 * t;
              const int4 qh = *reinterpret_cast<const int4 *>(qhb + phi * 16);
              const int4 ql = *reinterpret_cast<const int4 *>(qlb + phi * 16);
              acc1 = __dp4a(qh.x, wlo[t].x, acc1);
              acc1 = __dp4a(qh.y, wlo[t].y, acc1);
              acc1 = __dp4a(qh.z, wlo[t].z, acc1);
              acc1 = __dp4a(qh.w, wlo[t].w, acc1);
              acc2 = __dp4a(ql.x, whi[t].x, acc2);
              acc2 = __dp4a(ql.y, whi[t].y, acc2);
              acc2 = __dp4a(ql.z, whi[t].z, acc2);
              acc2 = __dp4a(ql.w, whi[t].w, acc2);
            }
            acc1 += __shfl_xor_sync(kFull, acc1, 8);
            acc2 += __shfl_xor_sync(kFull, acc2, 8);
            acc1 += __shfl_xor_sync(kFull, acc1, 16);
            acc2 += __shfl_xor_sync(kFull, acc2, 16);
            if (mine && part == 0) {
              const float4 m0 = *reinterpret_cast<const float4 *>(sQm + qr * QM);          // tmid, htight, qn, qlon
              const float4 m1 = *reinterpret_cast<const float4 *>(sQm + qr * QM + 4);      // sqh, sql, qA1, qB1
              const float2 m2 = *reinterpret_cast<const float2 *>(sQm + qr * QM + 8);      // qA2, qB2
              const float s2 = __uint_as_float(en.x) + (float)acc1 * (m1.x * em.x) + (float)acc2 * (m1.y * em.z);
              float tol = __fmaf_ru(m1.z, em.y, __fmul_ru(m1.w, em.x));     // int8 error of q_hi . e_lo
              tol = __fmaf_ru(m2.x, em.w, __fmaf_ru(m2.y, em.z, tol));      //              of q_lo . e_hi
              tol = __fadd_ru(__fmul_ru(tol, 1.0001f), m0.y);               // + the tight band (fp32 roundings of s2: 1e-4 of tol)
              const int pre = (en.x >> 31) == 0u ? 1 : 0;                   // the scan counted it as "above"
              int delta = -pre;
              if (s2 > tol) delta += 1;
              else if (s2 >= -tol) push_global(a, qt * QT + qr, (int)(a.shard_base + entity_of(a, erow)));
              if (delta) atomicAdd(a.cnt_gt + (int64_t)qt * QT + qr, delta);
            }
          }
          __syncwarp();   // the next tile's entries overwrite the list
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctrl->a_empty);   // this warp no longer reads the resident int8 rows
      if (q < a.Q && cnt) atomicAdd(a.cnt_gt + q, cnt);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();   // the leader's MMAs read the peer's shared memory until the very end
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<CG>(tmem, 512u);
  }
}

// ---- operand preparation ------------------------------------------------------------------

// fp32 rows (already in packed order) -> int8 copies of the fp16 lo and hi parts with one scale each
// per row, exactly the split skge_rank_pack_f16 stores (v = x * scale, hi = half(v), lo = half(v - hi)).
// One warp per row; rows beyond `rows` (padding up to rows_padded) are zero.
__global__ void __launch_bounds__(256) quant_rows_kernel(const float *__restrict__ X, int64_t rows,
                                                         int64_t rows_padded, int d, int kb, float scale,
                                                         int8_t *__restrict__ E8, float4 *__restrict__ meta,
                                                         float2 *__restrict__ norms) {
  const int lane = threadIdx.x & 31;
  int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = warp; r < rows_padded; r += nwarps) {
    float h[8], l[8];
    float mh = 0.f, ml = 0.f, l1h = 0.f, l1l = 0.f, n2h = 0.f, n2l = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = lane + 32 * i;
      const float v = (r < rows && k < d) ? __ldg(X + r * d + k) * scale : 0.f;
      const __half hh = __float2half_rn(v);
      h[i] = __half2float(hh);
      l[i] = __half2float(__float2half_rn(v - h[i]));
      mh = fmaxf(mh, fabsf(h[i]));
      ml = fmaxf(ml, fabsf(l[i]));
      l1h += fabsf(h[i]);
      l1l += fabsf(l[i]);
      n2h = fmaf(h[i], h[i], n2h);
      n2l = fmaf(l[i], l[i], n2l);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mh = fmaxf(mh, __shfl_xor_sync(kFull, mh, o));
      ml = fmaxf(ml, __shfl_xor_sync(kFull, ml, o));
      l1h += __shfl_xor_sync(kFull, l1h, o);
      l1l += __shfl_xor_sync(kFull, l1l, o);
      n2h += __shfl_xor_sync(kFull, n2h, o);
      n2l += __shfl_xor_sync(kFull, n2l, o);
    }
    const float invh = mh > 0.f ? 127.f / mh : 0.f, invl = ml > 0.f ? 127.f / ml : 0.f;
    int8_t *dst = E8 + r * (2 * (int64_t)kb);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = lane + 32 * i;
      if (k < kb) {
        dst[k] = (int8_t)max(-127, min(127, __float2int_rn(l[i] * invl)));
        dst[kb + k] = (int8_t)max(-127, min(127, __float2int_rn(h[i] * invh)));
      }
    }
    if (lane == 0) {
      meta[r] = make_float4(ml / 127.f, l1l * 1.001f, mh / 127.f, l1h * 1.001f);
      norms[r] = make_float2(sqrtf(n2l) * 1.001f, sqrtf(n2h) * 1.001f);   // ||e_lo||, ||e_hi|| (rounded up)
    }
  }
}

// Queries: swizzled int8 tiles of the fp16 hi and lo parts (h = half(q32 * qscale), exactly what
// skge_rank_pack_f16 stores; l = half(q32 * qscale - h)) and the per-query constants of the epilogue.
__global__ void __launch_bounds__(256) pack_q8x2_kernel(const float *__restrict__ q32, const float *__restrict__ qscale,
                                                        const float *__restrict__ thr_lo,
                                                        const float *__restrict__ thr_hi, int64_t Q, int d, int kch,
                                                        int8_t *__restrict__ Q8h, int8_t *__restrict__ Q8l,
                                                        float *__restrict__ qmeta) {
  const int lane = threadIdx.x & 31;
  const int kb = kch * 64;
  const bool swz = (kch & 1) == 0;
  int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t q = warp; q < Q; q += nwarps) {
    const float s = qscale[q];
    float h[8], l[8];
    float mh = 0.f, ml = 0.f, l1h = 0.f, l1l = 0.f, n2 = 0.f, n2l = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = lane + 32 * i;
      const float v = (k < d) ? __ldg(q32 + q * d + k) * s : 0.f;
      h[i] = __half2float(__float2half_rn(v));
      l[i] = __half2float(__float2half_rn(v - h[i]));
      mh = fmaxf(mh, fabsf(h[i]));
      ml = fmaxf(ml, fabsf(l[i]));
      l1h += fabsf(h[i]);
      l1l += fabsf(l[i]);
      n2 = fmaf(v, v, n2);
      n2l = fmaf(l[i], l[i], n2l);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mh = fmaxf(mh, __shfl_xor_sync(kFull, mh, o));
      ml = fmaxf(ml, __shfl_xor_sync(kFull, ml, o));
      l1h += __shfl_xor_sync(kFull, l1h, o);
      l1l += __shfl_xor_sync(kFull, l1l, o);
      n2 += __shfl_xor_sync(kFull, n2, o);
      n2l += __shfl_xor_sync(kFull, n2l, o);
    }
    const float sqh = mh / 127.f, sql = ml / 127.f;
    const float invh = mh > 0.f ? 127.f / mh : 0.f, invl = ml > 0.f ? 127.f / ml : 0.f;
    const int64_t tile = q / QT;
    const int r = (int)(q % QT);
    int8_t *dh = Q8h + (tile * QT + r) * kb, *dl = Q8l + (tile * QT + r) * kb;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = lane + 32 * i;
      if (k < kb) {
        const int c = k >> 4;
        const int phi = swz ? ((c & ~7) | ((c ^ r) & 7)) : c;
        dh[phi * 16 + (k & 15)] = (int8_t)max(-127, min(127, __float2int_rn(h[i] * invh)));
        dl[phi * 16 + (k & 15)] = (int8_t)max(-127, min(127, __float2int_rn(l[i] * invl)));
      }
    }
    if (lane == 0) {
      const float tlo = thr_lo[q], thi = thr_hi[q];
      const float tmid = 0.5f * tlo + 0.5f * thi;
      float *mq = qmeta + q * QM;
      mq[0] = tmid;
      mq[1] = fmaxf(__fsub_ru(thi, tmid), __fsub_ru(tmid, tlo));   // [tlo, thi] is inside tmid -+ htight
      mq[2] = sqrtf(n2) * 1.01f;                                  // ||q|| in scaled units; 1 % covers the fp32 roundings
      mq[3] = sqrtf(n2l) * 1.01f;                                 // ||q_lo||
      mq[4] = sqh;
      mq[5] = sql;
      mq[6] = 0.505f * sqh;                                       // times ||e_lo||_1: error of the int8 q_hi row
      mq[7] = 0.505f * (l1h + 0.5f * kb * sqh);                   // times scale_lo[e]: error of the int8 e_lo row
      mq[8] = 0.505f * sql;                                       // times ||e_hi||_1: error of the int8 q_lo row
      mq[9] = 0.505f * (l1l + 0.5f * kb * sql);                   // times scale_hi[e]: error of the int8 e_hi row
#pragma unroll
      for (int i = 10; i < QM; ++i) mq[i] = 0.f;
    }
  }
}

static int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

static size_t single_smem_bytes(int kch, int nb) {
  return (size_t)kch * BLOCK_BYTES + (size_t)2 * QT * kch * 64 + (size_t)QT * QM * 4 + (size_t)nb * STAGE_BYTES +
         sizeof(Ctrl) + sizeof(Lists);
}

}  // namespace rs
}  // namespace skge

using namespace skge;
using namespace skge::rs;

extern "C" {

int skge_rank_quant_rows(const float *X, int64_t rows, int d, float scale, void *E8, void *meta, void *norms,
                         skge_stream_t stream) {
  SKGE_REQUIRE(X && E8 && meta && norms && rows > 0 && d > 0, "bad arguments");
  SKGE_REQUIRE(d <= MAX_KCH * 64, "d <= 256");
  const int kb = (d + 63) / 64 * 64;
  const int64_t rp = round_up(rows, ET);
  int64_t blocks = (rp + 7) / 8;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  quant_rows_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(X, rows, rp, d, kb, scale, static_cast<int8_t *>(E8),
                                                               static_cast<float4 *>(meta),
                                                               static_cast<float2 *>(norms));
  SKGE_LAUNCH_CHECK();
  return 0;
}

int skge_rank_pack_q8x2(const float *q32, const float *qscale, const float *thr_lo, const float *thr_hi, int64_t Q,
                        int d, void *Q8h, void *Q8l, float *qmeta, skge_stream_t stream) {
  SKGE_REQUIRE(q32 && qscale && thr_lo && thr_hi && Q8h && Q8l && qmeta && Q >= 0 && d > 0, "bad arguments");
  SKGE_REQUIRE(d <= MAX_KCH * 64, "d <= 256");
  if (Q == 0) return 0;
  const int kch = (d + 63) / 64;
  int64_t blocks = (Q + 7) / 8;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  pack_q8x2_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(q32, qscale, thr_lo, thr_hi, Q, d, kch,
                                                              static_cast<int8_t *>(Q8h), static_cast<int8_t *>(Q8l),
                                                              qmeta);
  SKGE_LAUNCH_CHECK();
  return 0;
}

int skge_rank_single_count(const void *Ehi, const void *E8, const void *e_meta, const void *tile_w,
                           const int32_t *perm, int64_t n_shard, int64_t shard_base, const void *Qhi,
                           const void *Q8h, const void *Q8l, const float *qmeta, int64_t Q, int d, int cta_group,
                           int32_t *cnt_gt, int32_t *cand_q, int32_t *cand_e, int64_t cand_cap,
                           unsigned long long *cand_count, skge_stream_t stream) {
  SKGE_REQUIRE(Ehi && E8 && e_meta && tile_w && Qhi && Q8h && Q8l && qmeta && cnt_gt && cand_q && cand_e && cand_count,
               "null argument");
  SKGE_REQUIRE(d > 0 && d <= MAX_KCH * 64, "the tcgen05 ranking kernel supports d <= 256");
  SKGE_REQUIRE(cta_group == 1 || cta_group == 2, "cta_group must be 1 or 2");
  SKGE_REQUIRE(n_shard >= 0 && Q >= 0, "bad sizes");
  if (Q == 0 || n_shard == 0) return 0;
  SingleArgs a;
  a.Ehi = static_cast<const uint8_t *>(Ehi);
  a.E8 = static_cast<const int8_t *>(E8);
  a.e_meta = static_cast<const float4 *>(e_meta);
  a.tile_w = static_cast<const float2 *>(tile_w);
  a.perm = perm;
  a.Qhi = static_cast<const uint8_t *>(Qhi);
  a.Q8h = static_cast<const int8_t *>(Q8h);
  a.Q8l = static_cast<const int8_t *>(Q8l);
  a.qmeta = qmeta;
  a.n_shard = n_shard;
  a.shard_base = shard_base;
  a.Q = Q;
  a.kch = (d + 63) / 64;
  a.qtiles = (int)((Q + QT - 1) / QT);
  a.qunits = cta_group == 2 ? (a.qtiles + 1) / 2 : a.qtiles;
  a.etiles = (int)((n_shard + ET - 1) / ET);
  a.cnt_gt = cnt_gt;
  a.cand_q = cand_q;
  a.cand_e = cand_e;
  a.cand_cap = cand_cap;
  a.cand_count = cand_count;
  // as many 8 KB stages as fit beside the resident query tile
  const size_t smem_max = 232448;
  int nb = MAX_NB;
  while (nb > 2 && single_smem_bytes(a.kch, nb) > smem_max) --nb;
  SKGE_REQUIRE(single_smem_bytes(a.kch, nb) <= smem_max, "shared memory plan does not fit");
  a.nb = nb;
  const size_t smem = single_smem_bytes(a.kch, nb);
  // Entity tiles are walked in slices that stay L2-resident (fp16 hi blocks + both int8 rows) while
  // every query tile sweeps them; more slices when there are too few query units to fill the machine.
  const int nunits = cta_group == 2 ? kNumSMs / 2 : kNumSMs;
  int tps = (48 << 20) / (ET * a.kch * 64 * 4);
  if (tps > a.etiles) tps = a.etiles;
  while (tps > 32 && (int64_t)a.qunits * ((a.etiles + tps - 1) / tps) < 8 * nunits) tps = (tps + 1) / 2;
  a.tiles_per_slice = tps;
  a.nslices = (a.etiles + tps - 1) / tps;
  const int64_t nitems = (int64_t)a.qunits * a.nslices;
  int units = nitems < nunits ? (int)nitems : nunits;
  void (*kern)(const SingleArgs) = nullptr;
#define SKGE_SINGLE(CGV, K) (cta_group == CGV && a.kch == K) kern = rank_single_kernel<CGV, K>
  if SKGE_SINGLE(1, 1); else if SKGE_SINGLE(1, 2); else if SKGE_SINGLE(1, 3); else if SKGE_SINGLE(1, 4);
  else if SKGE_SINGLE(2, 1); else if SKGE_SINGLE(2, 2); else if SKGE_SINGLE(2, 3); else if SKGE_SINGLE(2, 4);
#undef SKGE_SINGLE
  SKGE_REQUIRE(kern != nullptr, "no kernel for this shape");
  SKGE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cta_group == 2 ? 2 * units : units);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = as_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cta_group;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SKGE_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
  SKGE_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
