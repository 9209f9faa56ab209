// placeholder: replaced by the tcgen05 kernel
#include "common.cuh"
extern "C" {
size_t skge_rank_packed_bytes(int64_t rows, int d) { (void)rows; (void)d; return 0; }
int skge_rank_pack_f16(const float *X, int64_t rows, int d, const float *row_scale, float scalar_scale, void *hi, void *lo, skge_stream_t stream) { skge::set_error("not built"); return SKGE_EINVAL; }
int skge_rank_gemm_count(const void *Ehi, const void *Elo, int64_t n_shard, int64_t shard_base, const void *Qhi, const void *Qlo, int64_t Q, int d, int nsplit, const float *thr_lo, const float *thr_hi, int32_t *cnt_gt, int32_t *cand_q, int32_t *cand_e, int64_t cand_cap, unsigned long long *cand_count, skge_stream_t stream) { skge::set_error("not built"); return SKGE_EINVAL; }
}
