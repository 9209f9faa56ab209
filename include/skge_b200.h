/*
 * skge_b200.h -- C ABI of libskge_b200.so: hand-written sm_100a kernels for the
 * scikit-kge embedding-training and filtered-ranking hot path.
 *
 * The reference (unmeshvrije/scikit-kge) has no FFI: its extension points are
 * Python duck-typed hooks.  Every entry point below names the reference
 * function(s) it replaces (paths relative to the reference root); the Python
 * package scikit-kge_b200/skge binds them with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the library never allocates or frees caller-visible memory: scratch space
 *     is a caller-provided workspace sized by the matching *_workspace_bytes();
 *   - every call is asynchronous on `stream` (a cudaStream_t) and keeps no
 *     global mutable state;
 *   - every function returns 0 on success, a negative code on failure
 *     (-cudaError_t for CUDA errors, SKGE_E* below otherwise); the message is
 *     available from skge_last_error() (thread-local);
 *   - index arrays are int32; parameters, AdaGrad state and gradients are fp32,
 *     row-major: E[N][d], R[M][d], W[M][d][d];
 *   - triples travel as three SoA arrays s[], o[], p[] (the reference's tuples
 *     are (s, o, p): skge/util.py:104-110).
 */
#ifndef SKGE_B200_H
#define SKGE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *skge_stream_t; /* cudaStream_t */

#define SKGE_API __attribute__((visibility("default")))

#define SKGE_EINVAL (-10001)     /* bad argument */
#define SKGE_EWORKSPACE (-10002) /* workspace too small */
#define SKGE_ENODEVICE (-10003)  /* no sm_100 device: there is no CPU fallback */

enum { SKGE_OPT_SGD = 0, SKGE_OPT_ADAGRAD = 1 };                       /* skge/param.py:124-158 */
enum { SKGE_POST_NONE = 0, SKGE_POST_NORMALIZE = 1, SKGE_POST_NORMLESS1 = 2 }; /* skge/param.py:161-174 */
enum { SKGE_AF_LINEAR = 0, SKGE_AF_SIGMOID = 1, SKGE_AF_TANH = 2, SKGE_AF_RELU = 3 }; /* skge/actfun.py:13-57 */
enum { SKGE_MODEL_TRANSE = 0, SKGE_MODEL_HOLE = 1, SKGE_MODEL_RESCAL = 2 };
enum { SKGE_RANK_L1 = 0, SKGE_RANK_DOT = 1 };

/* ---- library ---------------------------------------------------------- */
SKGE_API int skge_version(void);
SKGE_API const char *skge_last_error(void);
/* 0 if the current device is sm_100; SKGE_ENODEVICE otherwise. */
SKGE_API int skge_check_device(void);

/* ---- scores: Model._scores(ss, ps, os) -------------------------------- */
/* skge/transe.py:25-46 : -sum|E[s]+R[p]-E[o]| (l1) or -sum(.)^2 (no sqrt). */
SKGE_API int skge_scores_transe(const float *E, const float *R, const int32_t *s, const int32_t *p,
                       const int32_t *o, int64_t n, int d, int l1, float *out, skge_stream_t stream);
/* skge/hole.py:19-20 : sum_k R[p]_k ccorr(E[s],E[o])_k. */
SKGE_API int skge_scores_hole(const float *E, const float *R, const int32_t *s, const int32_t *p,
                     const int32_t *o, int64_t n, int d, float *out, skge_stream_t stream);
/* skge/rescal.py:31-35 : E[s]^T W[p] E[o]. */
SKGE_API int skge_scores_rescal(const float *E, const float *W, const int32_t *s, const int32_t *p,
                       const int32_t *o, int64_t n, int d, float *out, skge_stream_t stream);

/* ---- pairwise-margin minibatch: Model._pairwise_gradients + _batch_step -- */
/*
 * A minibatch is P (positive, negative) pairs given as six index arrays.
 * `valid` (nullable, uint8[P]) masks pairs the sampler could not produce.
 * counts (device int32[4]) receives {nviolations, U_E, U_R, 0}.
 *
 * *_grads returns exactly the reference's dict {'E': (ge, eidx), 'R': (gr, ridx)}:
 * ge[U_E][d] / gr[U_R][d] are the per-row MEANS over occurrences among the
 * violating pairs and eidx / ridx are ascending (skge/util.py:53-101).  The
 * caller sizes ge/eidx for min(4P, N) rows and gr/ridx for min(2P, M) rows.
 * nviolations == 0 is the reference's "return None" (skge/transe.py:90-91).
 *
 * *_step fuses the gradient with the parameter update of
 * StochasticTrainer._batch_step (skge/base.py:1306-1316): rows of E then R are
 * updated in place from gradients of the PRE-update parameters, by SGD or
 * AdaGrad (skge/param.py:124-158), followed by the row post-hook.
 * nviol_accum (nullable, device int64) += nviolations; ent_violations
 * (nullable, int32[N]) += 1 per distinct entity of each violating pair
 * (skge/transe.py:78-83); upd_counts_* (nullable) += 1 per updated row
 * (skge/param.py:149-150).
 */
SKGE_API size_t skge_pair_workspace_bytes(int64_t P, int d, int rows_per_pair, int64_t N, int64_t M);

/* skge/transe.py:48-165 */
SKGE_API int skge_transe_pair_grads(const float *E, const float *R, const int32_t *sp, const int32_t *op,
                           const int32_t *pp, const int32_t *sn, const int32_t *on,
                           const int32_t *pn, const uint8_t *valid, int64_t P, int64_t N,
                           int64_t M, int d, int l1, float margin, float *pscores,
                           float *nscores, float *ge, int32_t *eidx, float *gr, int32_t *ridx,
                           int32_t *counts, int32_t *ent_violations, void *ws, size_t ws_bytes,
                           skge_stream_t stream);
SKGE_API int skge_transe_pair_step(float *E, float *R, float *p2E, float *p2R, const int32_t *sp,
                          const int32_t *op, const int32_t *pp, const int32_t *sn,
                          const int32_t *on, const int32_t *pn, const uint8_t *valid, int64_t P,
                          int64_t N, int64_t M, int d, int l1, float margin, int opt, float lr,
                          int postE, int postR, int32_t *counts, int64_t *nviol_accum,
                          int32_t *ent_violations, int32_t *upd_counts_E, int32_t *upd_counts_R,
                          void *ws, size_t ws_bytes, skge_stream_t stream);

/* skge/hole.py:44-100 (af: skge/actfun.py; rparam on R rows only, hole.py:83) */
SKGE_API int skge_hole_pair_grads(const float *E, const float *R, const int32_t *sp, const int32_t *op,
                         const int32_t *pp, const int32_t *sn, const int32_t *on,
                         const int32_t *pn, const uint8_t *valid, int64_t P, int64_t N, int64_t M,
                         int d, int af, float margin, float rparam, float *pscores,
                         float *nscores, float *ge, int32_t *eidx, float *gr, int32_t *ridx,
                         int32_t *counts, void *ws, size_t ws_bytes, skge_stream_t stream);
SKGE_API int skge_hole_pair_step(float *E, float *R, float *p2E, float *p2R, const int32_t *sp,
                        const int32_t *op, const int32_t *pp, const int32_t *sn,
                        const int32_t *on, const int32_t *pn, const uint8_t *valid, int64_t P,
                        int64_t N, int64_t M, int d, int af, float margin, float rparam, int opt,
                        float lr, int postE, int postR, int32_t *counts, int64_t *nviol_accum,
                        int32_t *upd_counts_E, int32_t *upd_counts_R, void *ws, size_t ws_bytes,
                        skge_stream_t stream);

/*
 * HolE in the frequency domain (fused path; even d in [32, 1024] with d / 2 = 2^a 3^b 5^c:
 * every power of two and e.g. 100, 150, 200, 300 -- numpy's FFT in skge/util.py:27,50 takes any length).  Ehat / Rhat hold
 * the packed spectra of the rows of E / R (d floats per row: slot 0 = (X_0, X_{d/2}), slot f =
 * (Re X_f, Im X_f)); skge_hole_spectra fills them, skge_hole_pair_step_spectral is
 * skge_hole_pair_step with every per-pair transform removed: scores by Parseval, gradient rows
 * summed as spectra, one inverse + one forward transform per UPDATED row, which also keeps
 * Ehat / Rhat current (replaces the numpy.fft calls of skge/util.py:27,50 on this path).
 */
SKGE_API int skge_hole_spectra(const float *X, int64_t rows, int d, float *Xhat, skge_stream_t stream);
SKGE_API int skge_hole_pair_step_spectral(float *E, float *R, float *Ehat, float *Rhat, float *p2E, float *p2R,
                                 const int32_t *sp, const int32_t *op, const int32_t *pp,
                                 const int32_t *sn, const int32_t *on, const int32_t *pn,
                                 const uint8_t *valid, int64_t P, int64_t N, int64_t M, int d, int af,
                                 float margin, float rparam, int opt, float lr, int postE, int postR,
                                 int32_t *counts, int64_t *nviol_accum, int32_t *upd_counts_E,
                                 int32_t *upd_counts_R, void *ws, size_t ws_bytes, skge_stream_t stream);

/* ---- logistic minibatch: Model._gradients + _batch_step ---------------- */
/*
 * n labelled examples (s, o, p, y = +-1); `valid` (nullable, uint8[n]) masks examples the
 * sampler could not produce (the reference never appends those).  loss (device double,
 * nullable) is SET to sum logaddexp(0, -y*score) (skge/hole.py:26, skge/rescal.py:52);
 * loss_accum (nullable) += that value.  counts = {valid examples, U_E, U_second, 0}.
 */
SKGE_API size_t skge_logistic_workspace_bytes(int model, int64_t n, int d, int64_t N, int64_t M);

/* skge/hole.py:22-42 */
SKGE_API int skge_hole_logistic_grads(const float *E, const float *R, const int32_t *s, const int32_t *o,
                             const int32_t *p, const float *y, const uint8_t *valid, int64_t n, int64_t N,
                             int64_t M, int d, float rparam, float *ge, int32_t *eidx, float *gr,
                             int32_t *ridx, int32_t *counts, double *loss, void *ws,
                             size_t ws_bytes, skge_stream_t stream);
SKGE_API int skge_hole_logistic_step(float *E, float *R, float *p2E, float *p2R, const int32_t *s,
                            const int32_t *o, const int32_t *p, const float *y, const uint8_t *valid,
                            int64_t n, int64_t N, int64_t M, int d, float rparam, int opt, float lr,
                            int postE, int postR, int32_t *counts, double *loss_accum,
                            int32_t *upd_counts_E, int32_t *upd_counts_R, void *ws,
                            size_t ws_bytes, skge_stream_t stream);
/* skge/rescal.py:37-76 ; gw[U_W][d][d], pidx ascending */
SKGE_API int skge_rescal_logistic_grads(const float *E, const float *W, const int32_t *s, const int32_t *o,
                               const int32_t *p, const float *y, const uint8_t *valid, int64_t n, int64_t N,
                               int64_t M, int d, float rparam, float *ge, int32_t *eidx, float *gw,
                               int32_t *pidx, int32_t *counts, double *loss, void *ws,
                               size_t ws_bytes, skge_stream_t stream);
SKGE_API int skge_rescal_logistic_step(float *E, float *W, float *p2E, float *p2W, const int32_t *s,
                              const int32_t *o, const int32_t *p, const float *y, const uint8_t *valid,
                              int64_t n, int64_t N, int64_t M, int d, float rparam, int opt, float lr,
                              int postE, int postW, int32_t *counts, double *loss_accum,
                              int32_t *upd_counts_E, int32_t *upd_counts_W, void *ws,
                              size_t ws_bytes, skge_stream_t stream);

/* ---- ParameterUpdate.__call__(gradient, idx): skge/param.py:108-158 ----- */
/* param[idx] is updated from g[U][rowlen] (idx unique), then post on those rows.
 * p2 is AdaGrad's accumulator (ignored for SGD). rowlen = d, or d*d for W. */
SKGE_API int skge_sparse_update(float *param, float *p2, const float *g, const int32_t *idx, int64_t U,
                       int64_t rowlen, int opt, float lr, int post, int32_t *upd_counts,
                       skge_stream_t stream);
/* normalize / normless1 on rows idx (skge/param.py:161-174); idx == NULL: rows 0..U-1. */
SKGE_API int skge_rows_post(float *param, const int32_t *idx, int64_t U, int64_t rowlen, int post,
                   skge_stream_t stream);

/* ---- negative sampling: skge/sample.py:10-46, 91-110 ------------------- */
/*
 * The training-triple set is an open-addressing hash table of packed 64-bit
 * (s, o, p) keys built on the device.  skge_sample_corrupt emits, for each of
 * B positives, n_per rounds x the modes in modes_mask (bit0: subject, bit1:
 * object, bit2: predicate; ascending) one (positive, negative) pair: up to
 * ntries Philox draws of the corrupted slot, rejected while the candidate is a
 * training triple (and, with lcwa != 0, while (s', p') was never seen:
 * skge/sample.py:103-110, needs sp_table).  Output pair index =
 * (b * n_per + r) * nmodes + m.  out_valid[i] = 0 where all tries failed (the
 * reference skips those, sample.py:22-24).  The Philox counter of pair i is
 * offset + *offset_dev + i (offset_dev nullable, device memory): a step captured in a CUDA
 * graph keeps drawing fresh numbers when the caller advances *offset_dev between replays.
 */
SKGE_API size_t skge_tripleset_bytes(int64_t T);
SKGE_API int skge_tripleset_build(void *table, size_t table_bytes, const int32_t *s, const int32_t *o,
                         const int32_t *p, int64_t T, int pair_keys_only, skge_stream_t stream);
SKGE_API int skge_tripleset_contains(const void *table, size_t table_bytes, const int32_t *s,
                            const int32_t *o, const int32_t *p, int64_t n, uint8_t *out,
                            skge_stream_t stream);
SKGE_API int skge_sample_corrupt(const void *table, size_t table_bytes, const void *sp_table,
                        size_t sp_table_bytes, const int32_t *s, const int32_t *o,
                        const int32_t *p, const int32_t *batch_idx, int64_t B, int n_per,
                        int modes_mask, int64_t N, int64_t M, int ntries, uint64_t seed,
                        uint64_t offset, const uint64_t *offset_dev, int32_t *out_sp, int32_t *out_op,
                        int32_t *out_pp, int32_t *out_sn, int32_t *out_on, int32_t *out_pn,
                        uint8_t *out_valid, skge_stream_t stream);

/* ---- filtered ranking: FilteredRankingEval.positions -------------------- */
/*
 * skge/base.py:913-1031 with the per-model scorers of skge/run_transe.py:13-29
 * and skge/run_hole.py:10-19.  A query j is (kind[j], given[j], rel[j],
 * target[j]): kind 0 ranks target as the object of (given, rel, ?), kind 1 as
 * the subject of (?, rel, given).  Every model reduces to one query vector per
 * query and one sweep over the entity table:
 *     TransE : score(e) = -sum_i |e_i - q_i|   q = E[s]+R[p]  or  E[o]-R[p]
 *     HolE   : score(e) =  sum_i  e_i * q_i    q = cconv(R[p],E[s]) or ccorr(R[p],E[o])
 *     RESCAL : score(e) =  sum_i  e_i * q_i    q = W[p]^T E[s]   or  W[p] E[o]
 * rank = 1 + #{e : score(e) > score(target)} (equal to the reference's argsort
 * position whenever the target's score is not tied).
 *
 * skge_rank_make_queries writes q64 (exact, fp64), q32 (its fp32 rounding, for
 * the coarse sweep), tscore (exact fp64 target score), qnorm (||q||_2) and eps,
 * a bound on the coarse sweep's error: coarse_rel * ||q||_2 * enorm_max for the
 * DOT models (enorm_max >= max row norm of E), coarse_rel * (|tscore| +
 * ||q||_1 / (d+2)) for L1.  The coarse sweeps count entities whose score exceeds
 * tscore by more than eps and append the undecided (query, entity) pairs
 * (|score - tscore| <= eps) to a candidate list; skge_rank_rescore settles
 * candidates and filter entries in fp64 from the fp32 master table, so shard
 * boundaries never change a result.
 */
SKGE_API int skge_rank_make_queries(int model, const float *E, const float *RW, const uint8_t *kind,
                           const int32_t *given, const int32_t *rel, const int32_t *target,
                           int64_t Q, int d, float enorm_max, float coarse_rel, double *q64,
                           float *q32, double *tscore, float *eps, float *qnorm,
                           skge_stream_t stream);
/* CUDA-core coarse sweep in fp32 (skge/run_transe.py:13-29 for TransE, always L1; dot models with
 * d > 256) over rows [0, n_shard) of a shard (global ids shard_base + row).  Both operands are first
 * packed by skge_rank_sweep_pack into k-major tiles [tile of 128 rows][chunk of 16 k][k][row]
 * (skge_rank_sweep_packed_floats(rows, d) floats, zero padded) so that a pipeline stage is one
 * contiguous bulk-TMA copy per operand; Epk holds the shard's rows, Qpk the fp32 query vectors.
 * cnt_gt[Q] += definite wins; candidates appended at *cand_count (device, atomically); entries
 * beyond cand_cap are dropped but still counted, so the caller can detect overflow. */
SKGE_API int64_t skge_rank_sweep_packed_floats(int64_t rows, int d);
SKGE_API int skge_rank_sweep_pack(const float *src, int64_t rows, int d, float *out, skge_stream_t stream);
SKGE_API int skge_rank_sweep_tiles(int op, const float *Epk, int64_t n_shard, int64_t shard_base, int d,
                          const float *Qpk, const double *tscore, const float *eps, int64_t Q,
                          int32_t *cnt_gt, int32_t *cand_q, int32_t *cand_e, int64_t cand_cap,
                          unsigned long long *cand_count, skge_stream_t stream);
/* fp64 settlement of (query, entity) pairs: cnt[q] += 1 where
 * score64(q, e) > tscore[q]; pairs with e == target[q] are ignored when
 * target != NULL.  npairs_dev (device, nullable) overrides npairs and is
 * clamped to it.  Efull is the full fp32 table (global entity ids). */
SKGE_API int skge_rank_rescore(int op, const float *Efull, int d, const double *q64, const double *tscore,
                      const int32_t *pair_q, const int32_t *pair_e, int64_t npairs,
                      const unsigned long long *npairs_dev, const int32_t *target, int32_t *cnt,
                      skge_stream_t stream);
/* N scores of one query against the whole table, in fp64 (evaluator hooks
 * scores_o / scores_s of skge/run_transe.py:20-29, skge/run_hole.py:15-19). */
SKGE_API int skge_rank_scores_one(int op, const float *E, int64_t N, int d, const double *q64, double *out,
                         skge_stream_t stream);

/* tcgen05 path for the DOT models (HolE / RESCAL): the entity shard and the
 * query block are re-laid out as fp16 hi/lo split tiles (UMMA K-major core
 * matrices, 128 rows x 64 k per tile) and contracted on the 5th-gen tensor
 * cores with fp32 accumulation in TMEM; the epilogue compares against the
 * per-query thresholds straight out of TMEM and never stores a score. */
SKGE_API size_t skge_rank_packed_bytes(int64_t rows, int d); /* bytes of ONE (hi or lo) packed array */
/* lo_rowmajor (nullable): a second, row-major copy of the lo parts, [rows padded to 128][64 * ceil(d / 64)]
 * halfs -- the entity-side `Elo` argument of skge_rank_gemm_count when nsplit == 2.
 * lo_norm2 (nullable, [rows], zeroed by the caller): += squared L2 norm of each row's lo part. */
SKGE_API int skge_rank_pack_f16(const float *X, int64_t rows, int d, const float *row_scale,
                       float scalar_scale, void *hi, void *lo, void *lo_rowmajor, float *lo_norm2,
                       skge_stream_t stream);
/* Row-major fp16 lo rows (lo_rowmajor of skge_rank_pack_f16, rows padded to 128, row length
 * 64 * ceil(d / 64)) -> 8-bit rows: scale[row] = max|lo| / 127, byte = rn(lo / scale) + 128. */
SKGE_API int skge_rank_quant_lo(const void *lo_rowmajor, int64_t rows, int d, void *lo8, float *scale,
                       skge_stream_t stream);
/* Per-query power-of-two scale (max|q| * qscale in [2^11, 2^12)) and the scaled
 * thresholds thr = (tscore -+ eps) * qscale * escale, rounded outwards. */
SKGE_API int skge_rank_query_scale(const float *q32, const double *tscore, const float *eps, int64_t Q, int d,
                          float escale, float *qscale, float *thr_lo, float *thr_hi,
                          skge_stream_t stream);
/* nsplit: 1 = hi*hi only (fp16 accuracy), 3 = hi*hi + hi*lo + lo*hi on the tensor cores,
 * 2 = q_hi*e_hi + q_lo*e_hi on the tensor cores ("refine" mode): the accumulator of entity tile t
 * (128 packed rows) is tested against the tight thresholds widened by qwidth[q] * tile_w[t], which
 * must bound ||q|| * max ||e_lo|| over the tile in scaled units; pairs inside the wide band get
 * q_hi . e_lo added in the epilogue before the tight test.  Elo must then be the row-major lo
 * array of skge_rank_pack_f16.  perm (nullable) maps packed shard row -> shard-local entity id, so
 * the caller may pack the shard in any order (e.g. by row norm, which makes tile_w tight);
 * candidates always carry shard_base + entity id.  With lo_scale != NULL, Elo is instead the 8-bit
 * row-major array of skge_rank_quant_lo (half the gather traffic) and q1w[q] >= 0.5 ||q_hi||_1 in
 * scaled units: the quantisation error of a pair is at most q1w[q] * lo_scale[row], and that
 * pair's tight band is widened by it.  qwidth, tile_w, perm, lo_scale, q1w may be NULL unless
 * nsplit == 2. */
SKGE_API int skge_rank_gemm_count(const void *Ehi, const void *Elo, int64_t n_shard, int64_t shard_base,
                         const void *Qhi, const void *Qlo, int64_t Q, int d, int nsplit,
                         const float *thr_lo, const float *thr_hi, const float *qwidth,
                         const float *tile_w, const int32_t *perm, const float *lo_scale, const float *q1w,
                         int32_t *cnt_gt,
                         int32_t *cand_q, int32_t *cand_e, int64_t cand_cap,
                         unsigned long long *cand_count, skge_stream_t stream);

/* ---- refine kernel, second generation (csrc/rank_refine.cu) ------------------------------------
 * Same contract as skge_rank_gemm_count with nsplit == 2 (two fp16 products on the tensor cores, the
 * third one added in the epilogue for the pairs inside the wide band), rebuilt around UMMA N = 256:
 * the refinement reads 8-bit copies of BOTH operands (exact integer dot products) instead of fp16
 * query rows, and cta_group = 2 pairs two CTAs on one 256-query x 256-entity MMA so that each SM
 * stages half of every entity tile.  Replaces the ranking loop of skge/base.py:950-1017 for
 * skge/run_hole.py:15-19 scores on large sweeps.
 *
 * skge_rank_quant_lo_s8: row-major fp16 lo rows (lo_rowmajor of skge_rank_pack_f16) -> int8 rows
 *   e8 = rn(lo / scale), scale = max|lo| / 127, and meta[row] = (scale, ||lo||_1 rounded up) as float2.
 * skge_rank_pack_q8: int8 copy of the queries' fp16 hi parts (h = half(q32 * qscale), q8 = rn(h / sq),
 *   sq = max|h| / 127) in tiles of 128 rows x 64 * ceil(d / 64) bytes whose 16-byte chunks are
 *   XOR-swizzled with the row (bank-conflict-free gathers), and qmeta[q][8] = thr_lo, thr_hi,
 *   1.01 * qnorm * qscale, sq, qA, qB, 0, 0: the int8 arithmetic of pair (q, e) is off by at most
 *   qA[q] * meta[e].y + qB[q] * meta[e].x, which widens that pair's tight band.
 * skge_rank_refine_count: Ehi must hold an EVEN number of 128-row tiles (zero padded), tile_w one
 *   entry per 128-row tile (>= max ||e_lo||_2 over the tile, scaled units). */
SKGE_API int skge_rank_quant_lo_s8(const void *lo_rowmajor, int64_t rows, int d, void *lo8, void *meta,
                          skge_stream_t stream);
SKGE_API int skge_rank_pack_q8(const float *q32, const float *qscale, const float *qnorm, const float *thr_lo,
                      const float *thr_hi, int64_t Q, int d, void *Q8, float *qmeta, skge_stream_t stream);
SKGE_API int skge_rank_refine_count(const void *Ehi, const void *Elo8, const void *lo_meta, const float *tile_w,
                           const int32_t *perm, int64_t n_shard, int64_t shard_base, const void *Qhi,
                           const void *Qlo, const void *Q8, const float *qmeta, int64_t Q, int d, int cta_group,
                           int32_t *cnt_gt, int32_t *cand_q, int32_t *cand_e, int64_t cand_cap,
                           unsigned long long *cand_count, skge_stream_t stream);

/* ---- single-product kernel (csrc/rank_single.cu) -------------------------------------------------
 * Same contract again with ONE fp16 product on the tensor cores (q_hi . e_hi); both cross terms
 * (q_hi . e_lo and q_lo . e_hi) are added in the epilogue, from int8 copies of all four vectors, for
 * the pairs inside the wide band.  The accumulators are pre-loaded with minus the middle of each
 * query's undecided band, so the scan is a sign count plus a running |.| minimum.  Replaces the
 * ranking loop of skge/base.py:950-1017 for skge/run_hole.py:15-19 scores on large sweeps.
 *
 * skge_rank_quant_rows: fp32 rows (packed order) -> E8[row][2][kb] = int8 of the fp16 lo part, then of the
 *   fp16 hi part (the split skge_rank_pack_f16 stores for scalar_scale = scale), kb = 64 * ceil(d / 64);
 *   meta[row] = (scale_lo, ||lo||_1, scale_hi, ||hi||_1) as float4, norms[row] = (||lo||_2, ||hi||_2) as
 *   float2 (rounded up).  All three arrays have rows padded to a multiple of 256 (zeros).
 * skge_rank_pack_q8x2: swizzled int8 tiles of the queries' hi and lo parts (layout of skge_rank_pack_q8)
 *   and qmeta[q][16] = tmid, htight, ||q||, ||q_lo||, sq_hi, sq_lo, qA1, qB1, qA2, qB2, 0...; qmeta must be
 *   allocated for a multiple of 128 queries.
 * skge_rank_single_count: tile_w[128-row tile] = (max ||e_lo||_2, max (||e_hi||_2 + ||e_lo||_2)) as float2. */
SKGE_API int skge_rank_quant_rows(const float *X, int64_t rows, int d, float scale, void *E8, void *meta, void *norms,
                         skge_stream_t stream);
SKGE_API int skge_rank_pack_q8x2(const float *q32, const float *qscale, const float *thr_lo, const float *thr_hi,
                        int64_t Q, int d, void *Q8h, void *Q8l, float *qmeta, skge_stream_t stream);
SKGE_API int skge_rank_single_count(const void *Ehi, const void *E8, const void *e_meta, const void *tile_w,
                           const int32_t *perm, int64_t n_shard, int64_t shard_base, const void *Qhi,
                           const void *Q8h, const void *Q8l, const float *qmeta, int64_t Q, int d, int cta_group,
                           int32_t *cnt_gt, int32_t *cand_q, int32_t *cand_e, int64_t cand_cap,
                           unsigned long long *cand_count, skge_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SKGE_B200_H */
