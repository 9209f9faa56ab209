// HolE / RESCAL filtered ranking, large sweeps: two fp16 products on the tensor cores + an int8
// refinement of the third one in the epilogue  (skge/run_hole.py:15-19 as one GEMM whose result
// is never stored; rank definition skge/base.py:950-980, 994-1017).
//
//   score[q][e] = q_hi.e_hi + q_lo.e_hi   (tcgen05.mma, fp32 accumulators in TMEM)
//               + q_hi.e_lo               (only where it can change the outcome, see below)
//
// Every operand is split x * scale = hi + lo with hi, lo in fp16.  The missing product is bounded
// by ||q|| * max ||e_lo|| over the 128 packed rows of an entity tile (the caller packs the shard
// by decreasing row norm, so a tile's rows are alike); the epilogue first tests the accumulator
// against the thresholds widened by that bound.  The ~0.3 % of pairs inside the wide band get
//   q_hi.e_lo ~ sq[q] * se[e] * sum_k q8[q][k] * e8[e][k]
// added from 8-bit copies of both operands (dp4a, exact integer arithmetic; the quantisation
// error is bounded per pair by qA[q] * l1[e] + qB[q] * se[e], which widens that pair's tight
// band) and are then tested against the tight thresholds.  What stays undecided goes to the
// candidate list that skge_rank_rescore settles in fp64, so the result never depends on any of
// the low-precision arithmetic.
//
// Shape of the kernel (one persistent CTA per SM, 20 warps):
//   warp 0   one thread issues the bulk-TMA copies: a resident tile of 128 queries (fp16 hi and
//            lo UMMA blocks + the swizzled int8 rows), then the entity shard streamed in 8 KB stages
//   warp 1   one thread issues tcgen05.mma, UMMA N = 256: per MMA the tensor core reads 4 KB of A
//            and 8 KB of B from shared memory for 128 clocks of work (96 B/clk; the N = 128 shape of
//            rank_umma.cu needs 128 B/clk, all the shared-memory bandwidth there is)
//   warp 2   (cta_group::2 only, peer CTA) relays "my half of the stage has landed" to the leader
//   warps 4-19  epilogue: TMEM -> registers -> two compares per element folded into bit masks ->
//            popc; wide-band pairs -> per-warp list -> int8 refinement, four lanes per pair
// CG = 2 pairs two CTAs (cta_group::2, UMMA M = 256): each CTA keeps its own 128 queries and
// stages only half of every entity tile, which halves both the L2 -> SM fill traffic and the B
// operand reads per SM.
//
// Layouts (skge_rank_pack_f16 / skge_rank_pack_q8 / skge_rank_quant_lo_s8):
//   fp16 blocks  [tile of 128 rows][k chunk of 64] -> 16 KB, UMMA K-major no-swizzle core matrices
//                [kcore 8][rowgroup 16][row 8][8 halfs]
//   Q8           [query tile][128 rows][kb = 64 * kch bytes]; when kch is even the 16-byte chunk c
//                of row r sits at (c & ~7) | ((c ^ r) & 7), so that the eight lanes of a
//                quarter-warp, which read eight different rows, can pick chunks in eight different
//                bank groups
//   Elo8         [packed row][kb] int8, lo_meta[row] = (scale, ||e_lo||_1)
#include <cuda_fp16.h>

#include "common.cuh"
#include "umma.cuh"

namespace skge {
namespace rr {

using namespace ptx;

#ifndef SKGE_EPI_SLEEP
#define SKGE_EPI_SLEEP 128u
#endif
static constexpr int QT = 128;             // query rows per CTA (UMMA M per CTA)
static constexpr int ET = 256;             // entity rows per MMA (UMMA N)
static constexpr int BLOCK_BYTES = 16384;  // one (128-row tile, 64-k chunk) fp16 block
static constexpr int STAGE_BYTES = 8192;   // B stage: 256 rows x 16 k (CG = 1) or 128 rows x 32 k per CTA (CG = 2)
static constexpr int MAX_KCH = 4;          // d <= 256
static constexpr int MAX_NB = 12;
static constexpr int LIST_CAP = 32;        // wide-band pairs per warp and tile half
static constexpr int THREADS = 640;

struct __align__(8) Ctrl {
  uint64_t a_full, a_peer, a_empty;
  uint64_t b_full[MAX_NB], b_peer[MAX_NB], b_empty[MAX_NB];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_base, pad;
};
struct Lists {
  uint2 ent[16][LIST_CAP];   // (coarse score bits, lane << 8 | column)
  int count[16];
};

struct RefineArgs {
  const uint8_t *Ehi;        // fp16 blocks, an even number of 128-row tiles (zero padded)
  const int8_t *Elo8;        // [packed rows][kb]
  const float2 *lo_meta;     // [packed rows] (scale, l1)
  const float *tile_w;       // [2 * etiles] max ||e_lo||_2 per 128-row tile, rounded up
  const int32_t *perm;       // nullable: packed row -> shard-local entity id
  const uint8_t *Qhi, *Qlo;  // fp16 blocks [qtiles][kch]
  const int8_t *Q8;          // [qtiles][128][kb], swizzled
  const float *qmeta;        // [Q][8]: thr_lo, thr_hi, qwidth, sq, qA, qB, -, -
  int64_t n_shard, shard_base, Q;
  int kch, nb, qtiles, qunits, etiles, nslices, tiles_per_slice;
  int32_t *cnt_gt, *cand_q, *cand_e;
  int64_t cand_cap;
  unsigned long long *cand_count;
};

struct Item { int qunit, et_beg, et_end; };
__device__ __forceinline__ Item get_item(const RefineArgs &a, int item) {
  Item it;
  const int slice = item / a.qunits;
  it.qunit = item - slice * a.qunits;
  it.et_beg = slice * a.tiles_per_slice;
  it.et_end = min(a.etiles, it.et_beg + a.tiles_per_slice);
  return it;
}

__device__ __forceinline__ void push_global(const RefineArgs &a, int q, int e) {
  unsigned long long slot = atomicAdd(a.cand_count, 1ull);
  if ((int64_t)slot < a.cand_cap) {
    a.cand_q[slot] = q;
    a.cand_e[slot] = e;
  }
}
__device__ __forceinline__ int64_t entity_of(const RefineArgs &a, int64_t r) {
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

template <int CG, int KCH>
__global__ void __launch_bounds__(THREADS, 1) rank_refine_kernel(const RefineArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  constexpr int kch = KCH, kb = KCH * 64;   // compile-time row length: the refinement's addressing folds into immediates
  uint8_t *sA_hi = smem_raw;
  uint8_t *sA_lo = sA_hi + kch * BLOCK_BYTES;
  uint8_t *sQ8 = sA_lo + kch * BLOCK_BYTES;
  uint8_t *sB = sQ8 + QT * kb;
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

  // The three control warps run their loops warp-converged and elect one lane only around the
  // instruction that must be issued once: addresses and descriptors are then warp-uniform values
  // (uniform registers), not per-thread values that have to be broadcast before every UTCHMMA / UBLKCP.
  if (warp == 0) {
    // ===================== producer: bulk copies global -> shared =====================
    uint32_t bstage = 0, bphase = 0, aphase = 0;
    const uint32_t a_bytes = (uint32_t)(2 * kch * BLOCK_BYTES + QT * kb);
    for (int item = unit; item < nitems; item += nunits) {
      const Item it = get_item(a, item);
      const int qt = min(a.qtiles - 1, CG == 2 ? 2 * it.qunit + (int)rank : it.qunit);
      mbar_wait(&ctrl->a_empty, aphase ^ 1);  // previous item's MMAs retired, epilogue done with the int8 rows
      if (elect_one()) {
        mbar_expect_tx(&ctrl->a_full, a_bytes);
        const uint8_t *qh = a.Qhi + (int64_t)qt * kch * BLOCK_BYTES;
        const uint8_t *ql = a.Qlo + (int64_t)qt * kch * BLOCK_BYTES;
        for (int c = 0; c < kch; ++c) {
          bulk_g2s(sA_hi + c * BLOCK_BYTES, qh + (int64_t)c * BLOCK_BYTES, BLOCK_BYTES, &ctrl->a_full);
          bulk_g2s(sA_lo + c * BLOCK_BYTES, ql + (int64_t)c * BLOCK_BYTES, BLOCK_BYTES, &ctrl->a_full);
        }
        bulk_g2s(sQ8, a.Q8 + (int64_t)qt * QT * kb, (uint32_t)(QT * kb), &ctrl->a_full);
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
      // descriptors: only the start-address field changes (smem offsets < 256 KB never carry out of it)
      const uint64_t ahi0 = make_desc(smem_u32(sA_hi), 2048u, 128u);
      const uint64_t alo0 = make_desc(smem_u32(sA_lo), 2048u, 128u);
      const uint64_t b00 = make_desc(smem_u32(sB), B_LBO, 128u);
      uint32_t bstage = 0, bphase = 0, aphase = 0, accs = 0, accphase = 0;
      for (int item = unit; item < nitems; item += nunits) {
        const Item it = get_item(a, item);
        mbar_wait(&ctrl->a_full, aphase);
        if (CG == 2) mbar_wait(&ctrl->a_peer, aphase);
        aphase ^= 1;
        for (int et = it.et_beg; et < it.et_end; ++et) {
          mbar_wait(&ctrl->acc_empty[accs], accphase ^ 1);
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
                umma_f16<CG>(d_tmem, ahi0 + aoff, bd + (uint64_t)(j * 256), IDESC, (uint32_t)(ks | j));
                umma_f16<CG>(d_tmem, alo0 + aoff, bd + (uint64_t)(j * 256), IDESC, 1u);
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
    uint32_t tseq = 0, aphase = 0;
    for (int item = unit; item < nitems; item += nunits) {
      const Item it = get_item(a, item);
      const int qt = CG == 2 ? 2 * it.qunit + (int)rank : it.qunit;
      const int64_t q = (int64_t)qt * QT + row;
      float thi = INFINITY, tlo = INFINITY, qw = 0.f, sq = 0.f, qA = 0.f, qB = 0.f;
      if (q < a.Q) {
        const float4 m0 = __ldg(reinterpret_cast<const float4 *>(a.qmeta + q * 8));
        const float4 m1 = __ldg(reinterpret_cast<const float4 *>(a.qmeta + q * 8 + 4));
        tlo = m0.x; thi = m0.y; qw = m0.z; sq = m0.w;
        qA = m1.x; qB = m1.y;
      }
      mbar_wait_sleep<EPI_SLEEP>(&ctrl->a_full, aphase);   // the int8 query rows of this item are in shared memory
      aphase ^= 1;
      int cnt = 0;
      for (int et = it.et_beg; et < it.et_end; ++et) {
        const uint32_t my = tseq++;
        if ((my & 1u) != st) continue;                // the other quads' accumulator stage
        const int64_t e0 = (int64_t)et * ET + colhalf * 128;
        const int nvalid = (int)max((int64_t)0, min((int64_t)128, a.n_shard - e0));
        // this tile's wide band: the missing product is at most ||q|| * max ||e_lo|| over its rows
        const float tw = __ldg(a.tile_w + 2 * et + colhalf);
        const float whi = __fmaf_ru(qw, tw, thi), wlo = __fmaf_rd(-qw, tw, tlo);
        mbar_wait_sleep<EPI_SLEEP>(&ctrl->acc_full[st], (my >> 1) & 1u);
        tc_fence_after();
        const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + st * ET + colhalf * 128;
        int nlist = 0;
        for (int c2 = 0; c2 < 4; ++c2) {
          uint32_t r[32];
          tmem_ld32(taddr + 32 * c2, r);
          tmem_ld_wait();
          if (c2 == 3) {  // all 128 columns are in registers or consumed: hand the TMEM stage back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (CG == 2 && !leader) mbar_arrive_remote(&ctrl->acc_empty[st], 0);
              else mbar_arrive(&ctrl->acc_empty[st]);
            }
          }
          uint32_t mh[4] = {0u, 0u, 0u, 0u}, ml[4] = {0u, 0u, 0u, 0u};
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float v = __uint_as_float(r[j]);
            if (v > whi) mh[j & 3] |= 1u << j;
            if (v >= wlo) ml[j & 3] |= 1u << j;
          }
          uint32_t mhi = (mh[0] | mh[1]) | (mh[2] | mh[3]), mlo = (ml[0] | ml[1]) | (ml[2] | ml[3]);
          const int left = nvalid - c2 * 32;
          if (left < 32) {
            const uint32_t vm = left <= 0 ? 0u : (0xFFFFFFFFu >> (32 - left));
            mhi &= vm;
            mlo &= vm;
          }
          cnt += __popc(mhi);
          uint32_t band = mlo & ~mhi;   // inside the WIDE band: needs the missing product
          // -> the warp's list.  Slots come from a ballot (one element per lane, the usual case) or a warp
          // prefix sum of the per-lane counts: no shared-memory atomics in the dependent chain.
          const uint32_t hit = __ballot_sync(kFull, band != 0u);
          if (hit) {
            const int mine_n = __popc(band);
            int idx;
            if (__ballot_sync(kFull, mine_n > 1) == 0u) {
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
            while (band) {
              const int j = __ffs(band) - 1;
              band &= band - 1;
              if (idx < LIST_CAP) {
                wl->ent[w16][idx] = make_uint2(pick32(r, j), (uint32_t)((lane << 8) | (c2 * 32 + j)));
              } else {
                // list full (rare): let the fp64 pass settle this pair
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
            const int8_t *ebase = a.Elo8 + erow * kb;
            const uint8_t *qbase = sQ8 + qr * kb;
            const int x = swz ? (qr & 7) : 0;
            // every load of the pair is issued before the first use: the row's chunks and its constants
            int4 w[KCH];
#pragma unroll
            for (int t = 0; t < KCH; ++t) {
              const int c = swz ? (8 * (t >> 1) + (pi[t & 1] ^ x)) : part + 4 * t;
              w[t] = __ldg(reinterpret_cast<const int4 *>(ebase + c * 16));
            }
            const float2 meta = __ldg(a.lo_meta + erow);
            int acc = 0;
#pragma unroll
            for (int t = 0; t < KCH; ++t) {
              const int phi = swz ? 8 * (t >> 1) + pi[t & 1] : part + 4 * t;
              const int4 qv = *reinterpret_cast<const int4 *>(qbase + phi * 16);
              acc = __dp4a(qv.x, w[t].x, acc);
              acc = __dp4a(qv.y, w[t].y, acc);
              acc = __dp4a(qv.z, w[t].z, acc);
              acc = __dp4a(qv.w, w[t].w, acc);
            }
            acc += __shfl_xor_sync(kFull, acc, 8);
            acc += __shfl_xor_sync(kFull, acc, 16);
            // the thresholds live in the lane that owns the query row
            const float thi_p = __shfl_sync(kFull, thi, L), tlo_p = __shfl_sync(kFull, tlo, L);
            const float sq_p = __shfl_sync(kFull, sq, L);
            const float qA_p = __shfl_sync(kFull, qA, L), qB_p = __shfl_sync(kFull, qB, L);
            if (mine && part == 0) {
              const float s2 = __uint_as_float(en.x) + (float)acc * (sq_p * meta.x);
              const float tol = __fmaf_ru(qA_p, meta.y, __fmul_ru(qB_p, meta.x));
              const int64_t qg = (int64_t)qt * QT + qr;
              if (s2 > __fadd_ru(thi_p, tol)) atomicAdd(a.cnt_gt + qg, 1);
              else if (s2 >= __fadd_rd(tlo_p, -tol))
                push_global(a, (int)qg, (int)(a.shard_base + entity_of(a, erow)));
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

// Row-major fp16 lo rows -> int8 rows with one scale per row: scale = max|l| / 127,
// e8 = rn(l / scale); meta = (scale, ||l||_1 rounded up).  One warp per row.
__global__ void __launch_bounds__(256) quant_lo_s8_kernel(const __half *__restrict__ lo_rm, int64_t rows, int kb,
                                                          int8_t *__restrict__ lo8, float2 *__restrict__ meta) {
  const int lane = threadIdx.x & 31;
  int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = warp; r < rows; r += nwarps) {
    const __half *src = lo_rm + r * kb;
    float m = 0.f, l1 = 0.f;
    for (int k = lane; k < kb; k += 32) {
      const float x = fabsf(__half2float(src[k]));
      m = fmaxf(m, x);
      l1 += x;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      m = fmaxf(m, __shfl_xor_sync(kFull, m, o));
      l1 += __shfl_xor_sync(kFull, l1, o);
    }
    const float sc = m / 127.f, inv = m > 0.f ? 127.f / m : 0.f;
    for (int k = lane; k < kb; k += 32) {
      int i = __float2int_rn(__half2float(src[k]) * inv);
      i = max(-127, min(127, i));
      lo8[r * kb + k] = (int8_t)i;
    }
    if (lane == 0) meta[r] = make_float2(sc, l1 * 1.001f);
  }
}

// Queries: int8 copy of the fp16 hi parts (h = half(q32 * qscale), exactly what skge_rank_pack_f16
// stores) in the swizzled tile layout, and the per-query constants of the epilogue.
__global__ void __launch_bounds__(256) pack_q8_kernel(const float *__restrict__ q32, const float *__restrict__ qscale,
                                                      const float *__restrict__ qnorm,
                                                      const float *__restrict__ thr_lo,
                                                      const float *__restrict__ thr_hi, int64_t Q, int d, int kch,
                                                      int8_t *__restrict__ Q8, float *__restrict__ qmeta) {
  const int lane = threadIdx.x & 31;
  const int kb = kch * 64;
  const bool swz = (kch & 1) == 0;
  int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t q = warp; q < Q; q += nwarps) {
    const float s = qscale[q];
    float h[8];
    float m = 0.f, l1 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = lane + 32 * i;
      float v = (k < d) ? __ldg(q32 + q * d + k) * s : 0.f;
      h[i] = (k < kb) ? __half2float(__float2half_rn(v)) : 0.f;
      m = fmaxf(m, fabsf(h[i]));
      l1 += fabsf(h[i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      m = fmaxf(m, __shfl_xor_sync(kFull, m, o));
      l1 += __shfl_xor_sync(kFull, l1, o);
    }
    const float sq = m / 127.f, inv = m > 0.f ? 127.f / m : 0.f;
    const int64_t tile = q / QT;
    const int r = (int)(q % QT);
    int8_t *dst = Q8 + (tile * QT + r) * kb;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = lane + 32 * i;
      if (k < kb) {
        int v = __float2int_rn(h[i] * inv);
        v = max(-127, min(127, v));
        const int c = k >> 4;
        const int phi = swz ? ((c & ~7) | ((c ^ r) & 7)) : c;
        dst[phi * 16 + (k & 15)] = (int8_t)v;
      }
    }
    if (lane == 0) {
      float *mq = qmeta + q * 8;
      mq[0] = thr_lo[q];
      mq[1] = thr_hi[q];
      mq[2] = qnorm[q] * s * 1.01f;                 // ||q|| in scaled units; 1 % covers the fp32 roundings
      mq[3] = sq;
      mq[4] = 0.505f * sq;                          // times ||e_lo||_1: error of the int8 query row
      mq[5] = 0.505f * (l1 + 0.5f * kb * sq);       // times the entity row's scale: error of the int8 lo row
      mq[6] = 0.f;
      mq[7] = 0.f;
    }
  }
}

static int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

static size_t refine_smem_bytes(int kch, int nb) {
  return (size_t)2 * kch * BLOCK_BYTES + (size_t)QT * kch * 64 + (size_t)nb * STAGE_BYTES + sizeof(Ctrl) + sizeof(Lists);
}

}  // namespace rr
}  // namespace skge

using namespace skge;
using namespace skge::rr;

extern "C" {

int skge_rank_quant_lo_s8(const void *lo_rowmajor, int64_t rows, int d, void *lo8, void *meta,
                          skge_stream_t stream) {
  SKGE_REQUIRE(lo_rowmajor && lo8 && meta && rows > 0 && d > 0, "bad arguments");
  int kb = (d + 63) / 64 * 64;
  int64_t rp = round_up(rows, 128);
  int64_t blocks = (rp + 7) / 8;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  quant_lo_s8_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(static_cast<const __half *>(lo_rowmajor), rp, kb,
                                                                static_cast<int8_t *>(lo8),
                                                                static_cast<float2 *>(meta));
  SKGE_LAUNCH_CHECK();
  return 0;
}

int skge_rank_pack_q8(const float *q32, const float *qscale, const float *qnorm, const float *thr_lo,
                      const float *thr_hi, int64_t Q, int d, void *Q8, float *qmeta, skge_stream_t stream) {
  SKGE_REQUIRE(q32 && qscale && qnorm && thr_lo && thr_hi && Q8 && qmeta && Q >= 0 && d > 0, "bad arguments");
  SKGE_REQUIRE(d <= MAX_KCH * 64, "d <= 256");
  if (Q == 0) return 0;
  int kch = (d + 63) / 64;
  int64_t blocks = (Q + 7) / 8;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  pack_q8_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(q32, qscale, qnorm, thr_lo, thr_hi, Q, d, kch,
                                                            static_cast<int8_t *>(Q8), qmeta);
  SKGE_LAUNCH_CHECK();
  return 0;
}

int skge_rank_refine_count(const void *Ehi, const void *Elo8, const void *lo_meta, const float *tile_w,
                           const int32_t *perm, int64_t n_shard, int64_t shard_base, const void *Qhi,
                           const void *Qlo, const void *Q8, const float *qmeta, int64_t Q, int d, int cta_group,
                           int32_t *cnt_gt, int32_t *cand_q, int32_t *cand_e, int64_t cand_cap,
                           unsigned long long *cand_count, skge_stream_t stream) {
  SKGE_REQUIRE(Ehi && Elo8 && lo_meta && tile_w && Qhi && Qlo && Q8 && qmeta && cnt_gt && cand_q && cand_e &&
                   cand_count,
               "null argument");
  SKGE_REQUIRE(d > 0 && d <= MAX_KCH * 64, "the tcgen05 ranking kernel supports d <= 256");
  SKGE_REQUIRE(cta_group == 1 || cta_group == 2, "cta_group must be 1 or 2");
  SKGE_REQUIRE(n_shard >= 0 && Q >= 0, "bad sizes");
  if (Q == 0 || n_shard == 0) return 0;
  RefineArgs a;
  a.Ehi = static_cast<const uint8_t *>(Ehi);
  a.Elo8 = static_cast<const int8_t *>(Elo8);
  a.lo_meta = static_cast<const float2 *>(lo_meta);
  a.tile_w = tile_w;
  a.perm = perm;
  a.Qhi = static_cast<const uint8_t *>(Qhi);
  a.Qlo = static_cast<const uint8_t *>(Qlo);
  a.Q8 = static_cast<const int8_t *>(Q8);
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
  while (nb > 2 && refine_smem_bytes(a.kch, nb) > smem_max) --nb;
  SKGE_REQUIRE(refine_smem_bytes(a.kch, nb) <= smem_max, "shared memory plan does not fit");
  a.nb = nb;
  const size_t smem = refine_smem_bytes(a.kch, nb);
  // Entity tiles are walked in slices that stay L2-resident (fp16 hi blocks + int8 lo rows) while
  // every query tile sweeps them; more slices when there are too few query units to fill the machine.
  const int nunits = cta_group == 2 ? kNumSMs / 2 : kNumSMs;
  int tps = (48 << 20) / (ET * a.kch * 64 * 3);
  if (tps > a.etiles) tps = a.etiles;
  while (tps > 32 && (int64_t)a.qunits * ((a.etiles + tps - 1) / tps) < 8 * nunits) tps = (tps + 1) / 2;
  a.tiles_per_slice = tps;
  a.nslices = (a.etiles + tps - 1) / tps;
  const int64_t nitems = (int64_t)a.qunits * a.nslices;
  int units = nitems < nunits ? (int)nitems : nunits;
  void (*kern)(const RefineArgs) = nullptr;
#define SKGE_REFINE(CGV, K) (cta_group == CGV && a.kch == K) kern = rank_refine_kernel<CGV, K>
  if SKGE_REFINE(1, 1); else if SKGE_REFINE(1, 2); else if SKGE_REFINE(1, 3); else if SKGE_REFINE(1, 4);
  else if SKGE_REFINE(2, 1); else if SKGE_REFINE(2, 2); else if SKGE_REFINE(2, 3); else if SKGE_REFINE(2, 4);
#undef SKGE_REFINE
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
