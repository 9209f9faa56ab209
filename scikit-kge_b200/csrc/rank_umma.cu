// HolE / RESCAL filtered ranking on the 5th-generation tensor cores (tcgen05).
//
//   scores[q][e] = sum_k Q[q][k] * E[e][k]      (skge/run_hole.py:15-19 as one GEMM)
//
// is never materialised: a persistent, warp-specialised kernel streams the
// entity shard through shared memory with bulk-TMA copies, contracts it against
// a resident tile of 128 queries with tcgen05.mma (fp32 accumulators in TMEM) and
// the epilogue warps compare each accumulator, straight out of TMEM, with the
// query's two thresholds: above thr_hi -> counted, inside [thr_lo, thr_hi] ->
// appended to the candidate list that skge_rank_rescore settles in fp64.
//
// Precision: operands are fp16 hi/lo splits (x*scale = hi + lo + O(2^-22 |x|));
// with nsplit = 3 the three products hi*hi + hi*lo + lo*hi are accumulated into
// the same TMEM tile, which reproduces the fp32 inputs to ~2^-21 relative.  The
// thresholds already contain the (power-of-two) scales.
//
// nsplit = 2 ("refine" mode) issues only the two products that share the entity operand,
// q_hi*e_hi + q_lo*e_hi.  What is missing, q_hi . e_lo (+ the negligible q_lo . e_lo), is bounded by
// ||q|| max ||e_lo|| ~ 2^-13 ||q|| ||e|| (the maximum over the 128 rows of the entity tile; the
// caller packs the shard ordered by row norm so that a tile's rows are alike), so the epilogue first
// tests against thresholds widened by that bound; the ~0.3 % of pairs that fall into the wide band get the missing
// term added by their epilogue warp (32 lanes x 8 k each: q from the resident shared-memory
// tile, e_lo straight from the L2-resident slice) and are then tested against the tight
// thresholds exactly like an nsplit = 3 accumulator.  A third of the tensor-core work and half
// of the shared-memory fill traffic are gone; counts and candidates are identical in meaning.
// In this mode GemmArgs::Elo is the ROW-MAJOR copy of the lo parts ([rows padded to 128][kch * 64],
// skge_rank_pack_f16's optional third output): a gathered row is 4 contiguous lines.
//
// Memory layout (produced by skge_rank_pack_f16): rows are grouped in tiles of
// 128, k in chunks of 64; one (tile, chunk) block is 16 KB laid out as UMMA
// K-major, no-swizzle core matrices:  [kcore 8][rowgroup 16][row 8][8 halfs],
// i.e. leading-dimension byte offset (between the two k core matrices of one
// MMA) 2048 B and stride byte offset (between 8-row groups) 128 B.  A block is
// contiguous in global memory, so a stage is filled by plain 1-D bulk copies
// (no tensor map, nothing to keep in sync with a swizzle mode).
#include <cuda_fp16.h>

#include "common.cuh"

namespace skge {

static constexpr int TILE = 128;                       // rows per tile (queries: UMMA M, entities: UMMA N)
static constexpr int KCHUNK = 64;                      // k per block
static constexpr int BLOCK_HALFS = TILE * KCHUNK;      // 8192 halfs = 16 KB
static constexpr int BLOCK_BYTES = BLOCK_HALFS * 2;
static constexpr int MAX_KCH = 4;                      // d <= 256
static constexpr int B_STAGES = 3;                     // stages of (hi, lo) entity blocks
static constexpr int B_STAGES_REFINE = 5;              // refine mode: stages of hi blocks only
static constexpr int MAX_B_STAGES = 5;
static constexpr int LIST_CAP = 32;                    // refine mode: wide-band pairs per warp and tile
static constexpr int ACC_STAGES = 4;                   // 4 x 128 TMEM columns
static constexpr int STAGING = 192;                    // candidate staging entries in smem
static constexpr uint32_t LBO_BYTES = 2048, SBO_BYTES = 128;

// ---- PTX wrappers ------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// [0,14) start>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version = 1, [61,64) layout = 0.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(LBO_BYTES >> 4) << 16) |
         ((uint64_t)(SBO_BYTES >> 4) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor: D = F32 (bit 4), A = B = F16 (0), both K-major,
// N >> 3 at [17,23), M >> 4 at [24,29).
static constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(TILE >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- shared-memory plan ----------------------------------------------------------
struct __align__(8) Ctrl {
  uint64_t a_full, a_empty;
  uint64_t b_full[MAX_B_STAGES], b_empty[MAX_B_STAGES];
  uint64_t acc_full[ACC_STAGES], acc_empty[ACC_STAGES];
  uint32_t tmem_base;
  int stage_count;
  unsigned long long base_slot;
  int stage_q[STAGING], stage_e[STAGING];
};
// refine mode only (carved after Ctrl): per epilogue warp, the wide-band pairs of the current tile
struct WideLists {
  float score[16][LIST_CAP];   // coarse score
  int rc[16][LIST_CAP];        // lane << 8 | column in tile
  int count[16];
};

struct GemmArgs {
  const __half *Ehi, *Elo, *Qhi, *Qlo;
  int64_t n_shard, shard_base, Q;
  int kch, nsplit;
  const float *thr_lo, *thr_hi;
  // refine mode: the band is widened by qwidth[q] * tile_w[tile] >= ||q|| * max ||e_lo|| over the tile
  // (scaled units); perm (nullable) maps a packed shard row to its shard-local entity id
  const float *qwidth, *tile_w;
  const int32_t *perm;
  // refine mode with 8-bit lo rows (lo_scale != NULL): Elo points to uint8 rows [rows][kch * 64],
  // byte = round(e_lo / lo_scale[row]) + 128; q1w[q] >= 0.5 ||q_hi||_1 bounds the quantisation error
  // of one pair by q1w[q] * lo_scale[row], which widens that pair's band
  const float *lo_scale, *q1w;
  int32_t *cnt_gt, *cand_q, *cand_e;
  int64_t cand_cap;
  unsigned long long *cand_count;
  int qtiles, etiles;
  int nslices, tiles_per_slice;   // entity tiles are processed in L2-sized slices
};

// Work schedule shared by all roles.  The entity shard is cut into slices that fit in L2;
// a work item is (slice, query tile), numbered slice-major, and CTA c takes items
// c, c + grid, c + 2 grid, ...  All CTAs therefore sweep the SAME slice with different
// query tiles at the same time: a slice is fetched from HBM once and then served from L2
// to every query tile, and the fine-grained items keep the last wave short.
struct Item { int qt, et_beg, et_end; };
__device__ __forceinline__ Item get_item(const GemmArgs &a, int item) {
  Item it;
  int slice = item / a.qtiles;
  it.qt = item - slice * a.qtiles;
  it.et_beg = slice * a.tiles_per_slice;
  it.et_end = min(a.etiles, it.et_beg + a.tiles_per_slice);
  return it;
}

__device__ __forceinline__ void flush_staging(Ctrl *ctrl, const GemmArgs &a, int tid128) {
  // called by the 256 epilogue threads together
  asm volatile("bar.sync 1, 256;" ::: "memory");
  int n = ctrl->stage_count;
  if (n > STAGING) n = STAGING;
  if (tid128 == 0 && n > 0) ctrl->base_slot = atomicAdd(a.cand_count, (unsigned long long)n);
  asm volatile("bar.sync 1, 256;" ::: "memory");
  for (int i = tid128; i < n; i += 256) {
    unsigned long long slot = ctrl->base_slot + i;
    if ((int64_t)slot < a.cand_cap) {
      a.cand_q[slot] = ctrl->stage_q[i];
      a.cand_e[slot] = ctrl->stage_e[i];
    }
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");
  if (tid128 == 0) ctrl->stage_count = 0;
  asm volatile("bar.sync 1, 256;" ::: "memory");
}

// Append the in-band columns of one lane's 32-column chunk: one shared-memory atomic
// reserves the slots, then the set bits are walked.  Overflow of the staging area goes
// straight to the global list (rare).
__device__ __forceinline__ void push_band(Ctrl *ctrl, const GemmArgs &a, int q, int ebase, uint32_t band) {
  int n = __popc(band);
  int pos = atomicAdd(&ctrl->stage_count, n);
  while (band) {
    int j = __ffs(band) - 1;
    band &= band - 1;
    if (pos < STAGING) {
      ctrl->stage_q[pos] = q;
      ctrl->stage_e[pos] = ebase + j;
    } else {
      unsigned long long slot = atomicAdd(a.cand_count, 1ull);
      if ((int64_t)slot < a.cand_cap) {
        a.cand_q[slot] = q;
        a.cand_e[slot] = ebase + j;
      }
    }
    ++pos;
  }
}

__device__ __forceinline__ void push_one(Ctrl *ctrl, const GemmArgs &a, int q, int e) {
  int pos = atomicAdd(&ctrl->stage_count, 1);
  if (pos < STAGING) {
    ctrl->stage_q[pos] = q;
    ctrl->stage_e[pos] = e;
  } else {
    unsigned long long slot = atomicAdd(a.cand_count, 1ull);
    if ((int64_t)slot < a.cand_cap) {
      a.cand_q[slot] = q;
      a.cand_e[slot] = e;
    }
  }
}

// refine mode: the few pairs that stay undecided after the refinement go straight to the global
// list (no staging, so the epilogue warps never meet at a barrier)
__device__ __forceinline__ void push_global(const GemmArgs &a, int q, int e) {
  unsigned long long slot = atomicAdd(a.cand_count, 1ull);
  if ((int64_t)slot < a.cand_cap) {
    a.cand_q[slot] = q;
    a.cand_e[slot] = e;
  }
}

// shard-local entity id of packed row `r` (refine mode may pack the shard in another order)
__device__ __forceinline__ int64_t entity_of(const GemmArgs &a, int64_t r) { return a.perm ? (int64_t)__ldg(a.perm + r) : r; }

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

// sum over 8 k of q_hi * e_lo for one 16-byte chunk of each operand.  The term is a correction
// of relative size 2^-12, so packed fp16 arithmetic is enough (|q_hi| < 2^12 and |e_lo| <= 1 in
// scaled units: four products per half2 lane stay below 2^14, and the fp16 roundings contribute
// < 2^-21 ||q|| ||e||), and q_lo * e_lo (< 2^-23 ||q|| ||e||) is dropped like in the three-product
// mode.  Both are covered by the 2^-17 band.  Reading q_hi only also halves the shared-memory
// traffic of the refinement, which is what bounds it.
__device__ __forceinline__ float chunk_dot(const uint4 &qh, const uint4 &el) {
  const __half2 *h = reinterpret_cast<const __half2 *>(&qh);
  const __half2 *e = reinterpret_cast<const __half2 *>(&el);
  __half2 acc = __hmul2(h[0], e[0]);
  acc = __hfma2(h[1], e[1], acc);
  acc = __hfma2(h[2], e[2], acc);
  acc = __hfma2(h[3], e[3], acc);
  const float2 a = __half22float2(acc);
  return a.x + a.y;
}

// 16 k of q_hi (two 16-byte chunks) times 16 quantised e_lo bytes.  A byte b becomes the half
// 8 + b / 128 by dropping it into the mantissa of 0x4800 (= 8.0, ulp 2^-7); minus 9 gives
// (b - 128) / 128 exactly.  Four products per half2 accumulator lane (|q_hi| < 2^12, |x| <= 1).
__device__ __forceinline__ float dot16_q8(const uint4 &q0, const uint4 &q1, const uint4 &w) {
  const __half2 nine = __floats2half2_rn(9.f, 9.f);
  const __half2 *h0 = reinterpret_cast<const __half2 *>(&q0), *h1 = reinterpret_cast<const __half2 *>(&q1);
  const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
  __half2 A = __floats2half2_rn(0.f, 0.f), B = A;
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    uint32_t lo2 = __byte_perm(ww[m], 0x48484848u, 0x5140), hi2 = __byte_perm(ww[m], 0x48484848u, 0x5342);
    const __half2 x01 = __hsub2(*reinterpret_cast<__half2 *>(&lo2), nine);
    const __half2 x23 = __hsub2(*reinterpret_cast<__half2 *>(&hi2), nine);
    const __half2 *qq = m < 2 ? h0 : h1;
    A = __hfma2(x01, qq[2 * (m & 1)], A);
    B = __hfma2(x23, qq[2 * (m & 1) + 1], B);
  }
  const float2 fa = __half22float2(A), fb = __half22float2(B);
  return (fa.x + fa.y) + (fb.x + fb.y);
}

template <bool REFINE>
__global__ void __launch_bounds__(REFINE ? 640 : 384, 1) rank_gemm_kernel(GemmArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  // carve: A (kch blocks hi, kch blocks lo), B stages (hi, lo), control.  No-swizzle
  // descriptors and bulk copies only need 16-byte alignment.
  uint8_t *sA_hi = smem_raw;
  uint8_t *sA_lo = sA_hi + a.kch * BLOCK_BYTES;
  uint8_t *sB = sA_lo + a.kch * BLOCK_BYTES;  // stage s: hi at s*BSTRIDE, lo right after (not in refine mode)
  constexpr int NB = REFINE ? B_STAGES_REFINE : B_STAGES;
  constexpr int BSTRIDE = REFINE ? BLOCK_BYTES : 2 * BLOCK_BYTES;
  Ctrl *ctrl = reinterpret_cast<Ctrl *>(sB + NB * BSTRIDE);
  WideLists *wl = reinterpret_cast<WideLists *>(ctrl + 1);   // refine mode only

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kch = a.kch;
  const int nitems = a.qtiles * a.nslices;
  const bool use_lo = REFINE || a.nsplit == 3;   // q_lo resident / multiplied
  const bool use_b_lo = !REFINE && a.nsplit == 3; // e_lo streamed through shared memory

  if (threadIdx.x == 0) {
    mbar_init(&ctrl->a_full, 1);
    // refine mode: the epilogue warps read the query tile too, so they release it as well
    mbar_init(&ctrl->a_empty, REFINE ? 17 : 1);
    for (int s = 0; s < NB; ++s) { mbar_init(&ctrl->b_full[s], 1); mbar_init(&ctrl->b_empty[s], 1); }
    for (int s = 0; s < ACC_STAGES; ++s) { mbar_init(&ctrl->acc_full[s], 1); mbar_init(&ctrl->acc_empty[s], REFINE ? 4 : 8); }
    ctrl->stage_count = 0;
    if (REFINE) for (int w = 0; w < 16; ++w) wl->count[w] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {  // TMEM: all 512 columns (one CTA per SM)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctrl->tmem_base)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ctrl->tmem_base;

  if (warp == 0) {
    // ===================== producer: bulk copies global -> shared =====================
    if (lane == 0) {
      uint32_t bstage = 0, bphase = 0, aphase = 0;
      const uint32_t a_bytes = (uint32_t)kch * BLOCK_BYTES * (use_lo ? 2 : 1);
      const uint32_t b_bytes = (uint32_t)BLOCK_BYTES * (use_b_lo ? 2 : 1);
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const Item it = get_item(a, item);
        mbar_wait(&ctrl->a_empty, aphase ^ 1);  // previous item's MMAs retired
        mbar_expect_tx(&ctrl->a_full, a_bytes);
        const __half *qh = a.Qhi + (int64_t)it.qt * kch * BLOCK_HALFS;
        const __half *ql = a.Qlo + (int64_t)it.qt * kch * BLOCK_HALFS;
        for (int c = 0; c < kch; ++c) {
          bulk_g2s(sA_hi + c * BLOCK_BYTES, qh + (int64_t)c * BLOCK_HALFS, BLOCK_BYTES, &ctrl->a_full);
          if (use_lo) bulk_g2s(sA_lo + c * BLOCK_BYTES, ql + (int64_t)c * BLOCK_HALFS, BLOCK_BYTES, &ctrl->a_full);
        }
        aphase ^= 1;
        for (int et = it.et_beg; et < it.et_end; ++et) {
          for (int c = 0; c < kch; ++c) {
            mbar_wait(&ctrl->b_empty[bstage], bphase ^ 1);
            mbar_expect_tx(&ctrl->b_full[bstage], b_bytes);
            uint8_t *dst = sB + bstage * BSTRIDE;
            int64_t off = ((int64_t)et * kch + c) * BLOCK_HALFS;
            bulk_g2s(dst, a.Ehi + off, BLOCK_BYTES, &ctrl->b_full[bstage]);
            if (use_b_lo) bulk_g2s(dst + BLOCK_BYTES, a.Elo + off, BLOCK_BYTES, &ctrl->b_full[bstage]);
            if (++bstage == NB) { bstage = 0; bphase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one elected lane) =====================
    if (lane == 0) {
      uint32_t bstage = 0, bphase = 0, aphase = 0, accs = 0, accphase = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const Item it = get_item(a, item);
        mbar_wait(&ctrl->a_full, aphase);
        aphase ^= 1;
        for (int et = it.et_beg; et < it.et_end; ++et) {
          mbar_wait(&ctrl->acc_empty[accs], accphase ^ 1);  // epilogue drained this accumulator
          tc_fence_after();
          const uint32_t d_tmem = tmem + accs * TILE;
          uint32_t acc_on = 0;
          for (int c = 0; c < kch; ++c) {
            mbar_wait(&ctrl->b_full[bstage], bphase);
            tc_fence_after();
            const uint32_t a_hi = smem_u32(sA_hi + c * BLOCK_BYTES), a_lo = smem_u32(sA_lo + c * BLOCK_BYTES);
            const uint32_t b_hi = smem_u32(sB + bstage * BSTRIDE), b_lo = b_hi + BLOCK_BYTES;
#pragma unroll
            for (int ks = 0; ks < KCHUNK / 16; ++ks) {
              const uint32_t koff = ks * 2 * LBO_BYTES;  // one MMA consumes two k core matrices
              umma_f16(d_tmem, make_desc(a_hi + koff), make_desc(b_hi + koff), IDESC, acc_on);
              acc_on = 1;
              if (use_b_lo) umma_f16(d_tmem, make_desc(a_hi + koff), make_desc(b_lo + koff), IDESC, 1);
              if (use_lo) umma_f16(d_tmem, make_desc(a_lo + koff), make_desc(b_hi + koff), IDESC, 1);
            }
            tc_commit(&ctrl->b_empty[bstage]);  // frees the stage once these MMAs have read it
            if (++bstage == NB) { bstage = 0; bphase ^= 1; }
          }
          tc_commit(&ctrl->acc_full[accs]);  // accumulator complete -> epilogue
          if (++accs == ACC_STAGES) { accs = 0; accphase ^= 1; }
        }
        tc_commit(&ctrl->a_empty);  // all MMAs of this item retired -> A may be overwritten
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> compare -> count / candidates =====================
    const int quarter = warp & 3, cpart = (warp - 4) >> 2, w16 = warp - 4;   // cpart: column half / quad
    const int row = quarter * 32 + lane;  // query row inside the tile
    const int tid256 = threadIdx.x - 128;
    uint32_t accs = 0, accphase = 0, tseq = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
      const Item it = get_item(a, item);
      const int64_t q = (int64_t)it.qt * TILE + row;
      float thi = INFINITY, tlo = INFINITY;       // tight thresholds: t +- eps
      float qw = 0.f;                             // refine mode: ||q|| (scaled, rounded up)
      float q1 = 0.f;                             // refine mode, 8-bit lo rows: 0.5 ||q_hi||_1 (rounded up)
      if (q < a.Q) {
        thi = a.thr_hi[q];
        tlo = a.thr_lo[q];
        if (REFINE) {
          qw = a.qwidth[q];
          if (a.q1w) q1 = a.q1w[q];
        }
      }
      int cnt = 0;
      for (int et = it.et_beg; et < it.et_end; ++et) {
        const int64_t e0 = (int64_t)et * TILE;
        const int nvalid = (int)min((int64_t)TILE, a.n_shard - e0);
        if (!REFINE) {
          mbar_wait(&ctrl->acc_full[accs], accphase);
          tc_fence_after();
          const int col0 = cpart * 64;
          const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + accs * TILE + col0;
          uint32_t r[2][32];
          tmem_ld32(taddr, r[0]);
          tmem_ld32(taddr + 32, r[1]);
          tmem_ld_wait();
          // the accumulator is in registers: hand the TMEM stage back before the compare work
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ctrl->acc_empty[accs]);
          if (++accs == ACC_STAGES) { accs = 0; accphase ^= 1; }
#pragma unroll
          for (int c2 = 0; c2 < 2; ++c2) {
            // bit j of mhi / mlo: column j beats thr_hi / reaches thr_lo (4 partial masks keep
            // the dependency chains short)
            uint32_t mh[4] = {0u, 0u, 0u, 0u}, ml[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float v = __uint_as_float(r[c2][j]);
              if (v > thi) mh[j & 3] |= 1u << j;
              if (v >= tlo) ml[j & 3] |= 1u << j;
            }
            uint32_t mhi = (mh[0] | mh[1]) | (mh[2] | mh[3]), mlo = (ml[0] | ml[1]) | (ml[2] | ml[3]);
            const int left = nvalid - (col0 + c2 * 32);
            if (left < 32) {  // last, partial entity tile
              uint32_t vm = left <= 0 ? 0u : (0xFFFFFFFFu >> (32 - left));
              mhi &= vm;
              mlo &= vm;
            }
            cnt += __popc(mhi);
            uint32_t band = mlo & ~mhi;  // inside [thr_lo, thr_hi]: settle in fp64 later
            if (band) push_band(ctrl, a, (int)q, (int)(a.shard_base + e0 + col0 + c2 * 32), band);
          }
          // the flush decision must be uniform across the 256 epilogue threads: take it at fixed points
          if (((et - it.et_beg) & 15) == 15 || et == it.et_end - 1) flush_staging(ctrl, a, tid256);
        } else {
          const uint32_t my = tseq++;
          if ((int)(my & 3u) != cpart) continue;    // another quad's tile
          const uint32_t stage = my & 3u;
          // this tile's wide band: the missing product is at most ||q|| * max ||e_lo|| over its rows
          const float tw = __ldg(a.tile_w + et);
          const float whi = __fmaf_ru(qw, tw, thi), wlo = __fmaf_rd(-qw, tw, tlo);
          mbar_wait(&ctrl->acc_full[stage], (my >> 2) & 1u);
          tc_fence_after();
          const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + stage * TILE;
          for (int c2 = 0; c2 < 4; ++c2) {
            uint32_t r[32];
            tmem_ld32(taddr + 32 * c2, r);
            tmem_ld_wait();
            if (c2 == 3) {  // all 128 columns are in flight or consumed: hand the TMEM stage back
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&ctrl->acc_empty[stage]);
            }
            uint32_t mh[4] = {0u, 0u, 0u, 0u}, ml[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float v = __uint_as_float(r[j]);
              if (v > whi) mh[j & 3] |= 1u << j;
              if (v >= wlo) ml[j & 3] |= 1u << j;
            }
            uint32_t mhi = (mh[0] | mh[1]) | (mh[2] | mh[3]), mlo = (ml[0] | ml[1]) | (ml[2] | ml[3]);
            const int left = nvalid - c2 * 32;
            if (left < 32) {
              uint32_t vm = left <= 0 ? 0u : (0xFFFFFFFFu >> (32 - left));
              mhi &= vm;
              mlo &= vm;
            }
            cnt += __popc(mhi);
            uint32_t band = mlo & ~mhi;  // inside the WIDE band: needs the missing product
            // every lane walks its own bits and appends (lane, column, coarse score) to the warp's
            // list; the accumulator registers are read through a multiplexer
            while (band) {
              const int j = __ffs(band) - 1;
              band &= band - 1;
              const int idx = atomicAdd(&wl->count[w16], 1);
              if (idx < LIST_CAP) {
                wl->score[w16][idx] = __uint_as_float(pick32(r, j));
                wl->rc[w16][idx] = (lane << 8) | (c2 * 32 + j);
              } else {
                // list full (rare): let the fp64 pass settle this pair
                push_global(a, (int)q, (int)(a.shard_base + entity_of(a, e0 + c2 * 32 + j)));
              }
            }
          }
          __syncwarp();
          // Add q_hi . e_lo to the coarse score of every listed pair.  Eight pairs per
          // batch, four lanes per pair: lane 8 p + s takes the 16-byte k-chunks p, p + 4, ... of
          // pair s (chunk c = block c / 8, core matrix c % 8).  In the K-major core-matrix layout
          // all chunks of one query row live in the same four banks, and a 128-bit shared load is
          // served per quarter-warp: lanes 8 p .. 8 p + 7 belong to eight different pairs, so they
          // collide only where two rows agree modulo 8 (bucketing the list by row & 7 made the
          // loads conflict-free but doubled the number of half-empty batches: slower).
          const int slot = lane & 7, part = lane >> 3;
          const int n = min(wl->count[w16], LIST_CAP);
          if (n) {
            __syncwarp();
            if (lane == 0) wl->count[w16] = 0;
            for (int b = 0; b < n; b += 8) {
              const bool mine = b + slot < n;
              const float sc = mine ? wl->score[w16][b + slot] : 0.f;
              const int rc = mine ? wl->rc[w16][b + slot] : 0;
              const int qr = quarter * 32 + (rc >> 8), col = rc & 255;
              const uint32_t qbase = (qr >> 3) * SBO_BYTES + (qr & 7) * 16;
              float acc = 0.f, tol = 0.f;
              if (a.lo_scale) {
                // 8-bit lo rows: 256 B per pair instead of 512 B (the gathers bound this kernel)
                const uint8_t *ebase = reinterpret_cast<const uint8_t *>(a.Elo) + (e0 + col) * (int64_t)(kch * KCHUNK);
                const float sj = mine ? __ldg(a.lo_scale + e0 + col) : 0.f;
                uint4 w[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  const int cb = part + 4 * t;        // 16-byte chunk = 16 k
                  if (mine && cb < kch * 4) w[t] = __ldg(reinterpret_cast<const uint4 *>(ebase + cb * 16));
                }
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  const int cb = part + 4 * t;
                  if (mine && cb < kch * 4) {
                    const int ch = 2 * cb;            // q chunks ch, ch + 1 (8 k each, same 64-k block)
                    const uint32_t qoff = (ch >> 3) * BLOCK_BYTES + (ch & 7) * LBO_BYTES + qbase;
                    acc += dot16_q8(*reinterpret_cast<const uint4 *>(sA_hi + qoff),
                                    *reinterpret_cast<const uint4 *>(sA_hi + qoff + LBO_BYTES), w[t]);
                  }
                }
                acc *= 128.f * sj;
                tol = sj;   // times q1w of the owning lane, below
              } else {
                const __half *ebase = a.Elo + (e0 + col) * (int64_t)(kch * KCHUNK);   // row-major fp16 lo rows
                uint4 el[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                  const int ch = part + 4 * t;          // k-chunk 0 .. 31
                  if (mine && ch < kch * 8)
                    el[t] = __ldg(reinterpret_cast<const uint4 *>(ebase + ch * 8));
                }
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                  const int ch = part + 4 * t;
                  if (mine && ch < kch * 8) {
                    const uint32_t qoff = (ch >> 3) * BLOCK_BYTES + (ch & 7) * LBO_BYTES + qbase;
                    acc += chunk_dot(*reinterpret_cast<const uint4 *>(sA_hi + qoff), el[t]);
                  }
                }
              }
              acc += __shfl_xor_sync(kFull, acc, 8);
              acc += __shfl_xor_sync(kFull, acc, 16);
              const float s2 = sc + acc;
              // deliver each pair's refined score to the lane that owns its query row
              const uint32_t have = __ballot_sync(kFull, mine);
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const float su = __shfl_sync(kFull, s2, u);
                const int ru = __shfl_sync(kFull, rc, u);
                const float tu = __shfl_sync(kFull, tol, u) * q1;   // this pair's quantisation bound
                if (((have >> u) & 1u) && lane == (ru >> 8)) {
                  if (su > __fadd_ru(thi, tu)) ++cnt;
                  else if (su >= __fadd_rd(tlo, -tu)) push_global(a, (int)q, (int)(a.shard_base + entity_of(a, e0 + (ru & 255))));
                }
              }
            }
            __syncwarp();
          }
        }
      }
      if (REFINE) {  // this warp no longer reads the resident query tile
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctrl->a_empty);
      }
      if (q < a.Q && cnt) atomicAdd(a.cnt_gt + q, cnt);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// ---- operand packing ------------------------------------------------------------------
// X[rows][d] fp32 -> hi/lo fp16 blocks (see the layout at the top).  One thread per
// (row, group of 8 k): two 16-byte stores.
__global__ void __launch_bounds__(256) pack_f16_kernel(const float *__restrict__ X, int64_t rows, int d,
                                                       const float *__restrict__ row_scale, float scalar_scale,
                                                       __half *__restrict__ hi, __half *__restrict__ lo,
                                                       __half *__restrict__ lo_rm, float *__restrict__ lo_norm2,
                                                       int kch, int64_t rows_padded) {
  const int groups = kch * (KCHUNK / 8);
  int64_t total = rows_padded * groups;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = t / groups;
    int g = (int)(t - r * groups);
    int c = g / (KCHUNK / 8), kcore = g % (KCHUNK / 8);
    int k0 = c * KCHUNK + kcore * 8;
    float sc = scalar_scale * ((row_scale && r < rows) ? row_scale[r] : 1.0f);
    __align__(16) __half h[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = (r < rows && k0 + i < d) ? __ldg(X + r * d + k0 + i) * sc : 0.f;
      __half hh = __float2half_rn(v);
      h[i] = hh;
      l[i] = __float2half_rn(v - __half2float(hh));
    }
    int64_t tile = r / TILE;
    int rr = (int)(r % TILE);
    int64_t off = (tile * kch + c) * (int64_t)BLOCK_HALFS + (int64_t)kcore * (16 * 64) + (rr >> 3) * 64 + (rr & 7) * 8;
    *reinterpret_cast<uint4 *>(hi + off) = *reinterpret_cast<const uint4 *>(h);
    *reinterpret_cast<uint4 *>(lo + off) = *reinterpret_cast<const uint4 *>(l);
    // optional row-major copy of the lo part ([rows_padded][kch * 64]): what the refine-mode
    // epilogue gathers (a row's chunks are contiguous here, 2 KB apart in the blocked layout)
    if (lo_rm) *reinterpret_cast<uint4 *>(lo_rm + r * (int64_t)(kch * KCHUNK) + g * 8) = *reinterpret_cast<const uint4 *>(l);
    if (lo_norm2) {  // squared norm of the row's lo part (pre-zeroed by the caller): bounds the product refine mode defers
      float p2 = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float x = __half2float(l[i]);
        p2 = fmaf(x, x, p2);
      }
      if (r < rows && p2 != 0.f) atomicAdd(lo_norm2 + r, p2);
    }
  }
}

// Row-major fp16 lo rows -> 8-bit rows with one scale per row (refine mode's gather operand):
// scale = max|l| / 127, byte = rn(l / scale) + 128.  One warp per row.
__global__ void __launch_bounds__(256) quant_lo_kernel(const __half *__restrict__ lo_rm, int64_t rows, int kbytes,
                                                       uint8_t *__restrict__ lo8, float *__restrict__ scale) {
  const int lane = threadIdx.x & 31;
  int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = warp; r < rows; r += nwarps) {
    const __half *src = lo_rm + r * kbytes;
    float m = 0.f;
    for (int k = lane; k < kbytes; k += 32) m = fmaxf(m, fabsf(__half2float(src[k])));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(kFull, m, o));
    const float sc = m / 127.f, inv = m > 0.f ? 127.f / m : 0.f;
    for (int k = lane; k < kbytes; k += 32) {
      int i = __float2int_rn(__half2float(src[k]) * inv);
      i = max(-127, min(127, i));
      lo8[r * kbytes + k] = (uint8_t)(i + 128);
    }
    if (lane == 0) scale[r] = sc;
  }
}

// Per-query power-of-two scale and the scaled thresholds.
//   qscale = 2^(12 - ceil(log2 max|q|));  thr = (tscore -+ eps) * qscale * escale (rounded outwards)
__global__ void __launch_bounds__(256) query_scale_kernel(const float *__restrict__ q32,
                                                          const double *__restrict__ tscore,
                                                          const float *__restrict__ eps, int64_t Q, int d,
                                                          float escale, float *__restrict__ qscale,
                                                          float *__restrict__ thr_lo, float *__restrict__ thr_hi) {
  const int lane = threadIdx.x & 31;
  int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t q = warp; q < Q; q += nwarps) {
    float m = 0.f;
    for (int c = lane; c < d; c += 32) m = fmaxf(m, fabsf(__ldg(q32 + q * d + c)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(kFull, m, o));
    if (lane == 0) {
      int ex = 0;
      if (m > 0.f) frexpf(m, &ex);          // m = f * 2^ex, f in [0.5, 1)
      float s = ldexpf(1.0f, 12 - ex);       // m * s in [2^11, 2^12)
      qscale[q] = s;
      double f = (double)s * (double)escale, t = tscore[q], e = (double)eps[q];
      thr_hi[q] = __double2float_ru((t + e) * f);
      thr_lo[q] = __double2float_rd((t - e) * f);
    }
  }
}

static int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

}  // namespace skge

using namespace skge;

extern "C" {

size_t skge_rank_packed_bytes(int64_t rows, int d) {
  int kch = (d + KCHUNK - 1) / KCHUNK;
  return (size_t)(round_up(rows > 0 ? rows : 1, TILE) / TILE) * kch * BLOCK_BYTES;
}

int skge_rank_pack_f16(const float *X, int64_t rows, int d, const float *row_scale, float scalar_scale,
                       void *hi, void *lo, void *lo_rowmajor, float *lo_norm2, skge_stream_t stream) {
  SKGE_REQUIRE(X && hi && lo && rows > 0 && d > 0, "bad arguments");
  int kch = (d + KCHUNK - 1) / KCHUNK;
  int64_t rp = round_up(rows, TILE);
  int64_t total = rp * kch * (KCHUNK / 8);
  int64_t blocks = (total + 255) / 256;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  pack_f16_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(X, rows, d, row_scale, scalar_scale,
                                                             static_cast<__half *>(hi), static_cast<__half *>(lo),
                                                             static_cast<__half *>(lo_rowmajor), lo_norm2, kch, rp);
  SKGE_LAUNCH_CHECK();
  return 0;
}

int skge_rank_quant_lo(const void *lo_rowmajor, int64_t rows, int d, void *lo8, float *scale,
                       skge_stream_t stream) {
  SKGE_REQUIRE(lo_rowmajor && lo8 && scale && rows > 0 && d > 0, "bad arguments");
  int kbytes = (d + KCHUNK - 1) / KCHUNK * KCHUNK;
  int64_t rp = round_up(rows, TILE);
  int64_t blocks = (rp + 7) / 8;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  quant_lo_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(static_cast<const __half *>(lo_rowmajor), rp, kbytes,
                                                             static_cast<uint8_t *>(lo8), scale);
  SKGE_LAUNCH_CHECK();
  return 0;
}

int skge_rank_query_scale(const float *q32, const double *tscore, const float *eps, int64_t Q, int d,
                          float escale, float *qscale, float *thr_lo, float *thr_hi, skge_stream_t stream) {
  SKGE_REQUIRE(q32 && tscore && eps && qscale && thr_lo && thr_hi && Q >= 0 && d > 0, "bad arguments");
  if (Q == 0) return 0;
  int64_t blocks = (Q + 7) / 8;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  query_scale_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(q32, tscore, eps, Q, d, escale, qscale, thr_lo,
                                                                thr_hi);
  SKGE_LAUNCH_CHECK();
  return 0;
}

int skge_rank_gemm_count(const void *Ehi, const void *Elo, int64_t n_shard, int64_t shard_base,
                         const void *Qhi, const void *Qlo, int64_t Q, int d, int nsplit,
                         const float *thr_lo, const float *thr_hi, const float *qwidth,
                         const float *tile_w, const int32_t *perm, const float *lo_scale, const float *q1w,
                         int32_t *cnt_gt,
                         int32_t *cand_q, int32_t *cand_e, int64_t cand_cap,
                         unsigned long long *cand_count, skge_stream_t stream) {
  SKGE_REQUIRE(Ehi && Elo && Qhi && Qlo && thr_lo && thr_hi && cnt_gt && cand_q && cand_e && cand_count,
               "null argument");
  SKGE_REQUIRE(d > 0 && d <= MAX_KCH * KCHUNK, "the tcgen05 ranking kernel supports d <= 256");
  SKGE_REQUIRE(nsplit >= 1 && nsplit <= 3, "nsplit must be 1, 2 or 3");
  SKGE_REQUIRE(nsplit != 2 || (qwidth && tile_w), "nsplit = 2 needs qwidth and tile_w");
  SKGE_REQUIRE((lo_scale == nullptr) == (q1w == nullptr), "lo_scale and q1w go together");
  SKGE_REQUIRE(n_shard >= 0 && Q >= 0, "bad sizes");
  if (Q == 0 || n_shard == 0) return 0;
  GemmArgs a;
  a.Ehi = static_cast<const __half *>(Ehi);
  a.Elo = static_cast<const __half *>(Elo);
  a.Qhi = static_cast<const __half *>(Qhi);
  a.Qlo = static_cast<const __half *>(Qlo);
  a.n_shard = n_shard;
  a.shard_base = shard_base;
  a.Q = Q;
  a.kch = (d + KCHUNK - 1) / KCHUNK;
  a.nsplit = nsplit;
  a.thr_lo = thr_lo;
  a.thr_hi = thr_hi;
  a.qwidth = qwidth;
  a.tile_w = tile_w;
  a.perm = perm;
  a.lo_scale = lo_scale;
  a.q1w = q1w;
  a.cnt_gt = cnt_gt;
  a.cand_q = cand_q;
  a.cand_e = cand_e;
  a.cand_cap = cand_cap;
  a.cand_count = cand_count;
  a.qtiles = (int)((Q + TILE - 1) / TILE);
  a.etiles = (int)((n_shard + TILE - 1) / TILE);
  const bool refine = nsplit == 2;
  size_t smem = (size_t)2 * a.kch * BLOCK_BYTES +
                (refine ? (size_t)B_STAGES_REFINE * BLOCK_BYTES + sizeof(WideLists) : (size_t)B_STAGES * 2 * BLOCK_BYTES) +
                sizeof(Ctrl);
  if (refine)
    SKGE_CUDA(cudaFuncSetAttribute(rank_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else
    SKGE_CUDA(cudaFuncSetAttribute(rank_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // slices of at most ~48 MB (hi + lo) so a slice stays L2-resident while every query tile
  // sweeps it; more slices when there are too few query tiles to fill the machine
  int tps = (48 << 20) / (a.kch * 2 * BLOCK_BYTES);
  if (tps > a.etiles) tps = a.etiles;
  while (tps > 64 && (int64_t)a.qtiles * ((a.etiles + tps - 1) / tps) < 8 * kNumSMs) tps = (tps + 1) / 2;
  a.tiles_per_slice = tps;
  a.nslices = (a.etiles + tps - 1) / tps;
  int64_t nitems = (int64_t)a.qtiles * a.nslices;
  int grid = nitems < kNumSMs ? (int)nitems : kNumSMs;
  if (refine) rank_gemm_kernel<true><<<grid, 640, smem, as_stream(stream)>>>(a);
  else rank_gemm_kernel<false><<<grid, 384, smem, as_stream(stream)>>>(a);
  SKGE_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
