"""Thin tensor-level wrappers over the C ABI (one function per entry point family).

Everything here takes and returns CUDA tensors; nothing computes on the host.
"""
import torch

from . import _ext
from ._ext import check, lib, ptr, stream

_ws = _ext.Workspace()

# Number of libskge_b200 kernels launched so far (CUB sort/scan passes and memsets
# are not counted): bench.py reports the delta over its timed region.
LAUNCHES = {'n': 0}
_OWN_KERNELS = {'scores': 1, 'pair': 7, 'logistic_hole': 7, 'logistic_rescal': 13, 'sample': 1, 'set': 1,
                'make_queries': 1, 'sweep': 1, 'sweep_pack': 1, 'rescore': 1, 'scores_one': 1, 'pack': 1, 'gemm': 1}


def _count(what, times=1):
    LAUNCHES['n'] += _OWN_KERNELS[what] * times

TRANSE, HOLE, RESCAL = _ext.MODEL_TRANSE, _ext.MODEL_HOLE, _ext.MODEL_RESCAL


def _i32(n, dev=None):
    return torch.empty(n, dtype=torch.int32, device=dev or _ext.device())


def _f32(*shape):
    return torch.empty(*shape, dtype=torch.float32, device=_ext.device())


def scores(model, E, R, s, p, o, l1=True):
    """Model._scores: s, p, o are int32 CUDA tensors; returns fp32 [n]."""
    n, d = s.numel(), E.shape[1]
    out = _f32(n)
    _count('scores')
    if model == TRANSE:
        check(lib().skge_scores_transe(ptr(E), ptr(R), ptr(s), ptr(p), ptr(o), n, d, int(bool(l1)), ptr(out),
                                       stream()))
    elif model == HOLE:
        check(lib().skge_scores_hole(ptr(E), ptr(R), ptr(s), ptr(p), ptr(o), n, d, ptr(out), stream()))
    else:
        check(lib().skge_scores_rescal(ptr(E), ptr(R), ptr(s), ptr(p), ptr(o), n, d, ptr(out), stream()))
    return out


def pair_workspace(P, d, rows, N, M):
    return _ws.get(lib().skge_pair_workspace_bytes(P, d, rows, N, M))


def pair_grads(model, E, R, pos, neg, valid, margin, l1_or_af, rparam=0.0, ent_viol=None):
    """Un-fused pairwise gradients.  pos/neg are (s, o, p) triples of int32 CUDA
    tensors.  Returns dict(nviol, ge, eidx, gr, ridx, pscores, nscores)."""
    (sp, op, pp), (sn, on, pn) = pos, neg
    P, (N, d), M = sp.numel(), E.shape, R.shape[0]
    ue, ur = min(4 * P, N), min(2 * P, M)
    ge, eidx, gr, ridx = _f32(ue, d), _i32(ue), _f32(ur, d), _i32(ur)
    ps, ns, counts = _f32(P), _f32(P), _i32(4)
    rows = 2 if model == TRANSE else 6
    ws = pair_workspace(P, d, rows, N, M)
    _count('pair')
    if model == TRANSE:
        check(lib().skge_transe_pair_grads(ptr(E), ptr(R), ptr(sp), ptr(op), ptr(pp), ptr(sn), ptr(on), ptr(pn),
                                           ptr(valid), P, N, M, d, int(l1_or_af), float(margin), ptr(ps), ptr(ns),
                                           ptr(ge), ptr(eidx), ptr(gr), ptr(ridx), ptr(counts), ptr(ent_viol),
                                           ptr(ws), ws.numel(), stream()))
    else:
        check(lib().skge_hole_pair_grads(ptr(E), ptr(R), ptr(sp), ptr(op), ptr(pp), ptr(sn), ptr(on), ptr(pn),
                                         ptr(valid), P, N, M, d, int(l1_or_af), float(margin), float(rparam),
                                         ptr(ps), ptr(ns), ptr(ge), ptr(eidx), ptr(gr), ptr(ridx), ptr(counts),
                                         ptr(ws), ws.numel(), stream()))
    nviol, U_E, U_R, _ = counts.tolist()
    return dict(nviol=nviol, ge=ge[:U_E], eidx=eidx[:U_E], gr=gr[:U_R], ridx=ridx[:U_R], pscores=ps, nscores=ns)


def pair_step(model, E, R, p2E, p2R, pos, neg, valid, margin, l1_or_af, rparam, opt, lr, postE, postR,
              counts, nviol_accum, ent_viol=None, ucE=None, ucR=None):
    """Fused minibatch step (gradient + update), asynchronous on the stream."""
    (sp, op, pp), (sn, on, pn) = pos, neg
    P, (N, d), M = sp.numel(), E.shape, R.shape[0]
    rows = 2 if model == TRANSE else 6
    ws = pair_workspace(P, d, rows, N, M)
    _count('pair')
    if model == TRANSE:
        check(lib().skge_transe_pair_step(ptr(E), ptr(R), ptr(p2E), ptr(p2R), ptr(sp), ptr(op), ptr(pp), ptr(sn),
                                          ptr(on), ptr(pn), ptr(valid), P, N, M, d, int(l1_or_af), float(margin),
                                          opt, float(lr), postE, postR, ptr(counts), ptr(nviol_accum),
                                          ptr(ent_viol), ptr(ucE), ptr(ucR), ptr(ws), ws.numel(), stream()))
    else:
        check(lib().skge_hole_pair_step(ptr(E), ptr(R), ptr(p2E), ptr(p2R), ptr(sp), ptr(op), ptr(pp), ptr(sn),
                                        ptr(on), ptr(pn), ptr(valid), P, N, M, d, int(l1_or_af), float(margin),
                                        float(rparam), opt, float(lr), postE, postR, ptr(counts),
                                        ptr(nviol_accum), ptr(ucE), ptr(ucR), ptr(ws), ws.numel(), stream()))


def hole_spectra(X, out=None):
    """Packed spectra of the rows of X (even d in [32, 1024] with d / 2 = 2^a 3^b 5^c), see csrc/fft.cuh."""
    if out is None:
        out = torch.empty_like(X)
    _count('scores')
    check(lib().skge_hole_spectra(ptr(X), X.shape[0], X.shape[1], ptr(out), stream()))
    return out


def hole_pair_step_spectral(E, R, Ehat, Rhat, p2E, p2R, pos, neg, valid, margin, af, rparam, opt, lr, postE, postR,
                            counts, nviol_accum, ucE=None, ucR=None):
    (sp, op, pp), (sn, on, pn) = pos, neg
    P, (N, d), M = sp.numel(), E.shape, R.shape[0]
    ws = pair_workspace(P, d, 6, N, M)
    _count('pair')
    check(lib().skge_hole_pair_step_spectral(ptr(E), ptr(R), ptr(Ehat), ptr(Rhat), ptr(p2E), ptr(p2R), ptr(sp),
                                             ptr(op), ptr(pp), ptr(sn), ptr(on), ptr(pn), ptr(valid), P, N, M, d,
                                             int(af), float(margin), float(rparam), opt, float(lr), postE, postR,
                                             ptr(counts), ptr(nviol_accum), ptr(ucE), ptr(ucR), ptr(ws), ws.numel(),
                                             stream()))


def logistic_grads(model, E, R2, s, o, p, y, rparam, valid=None):
    """Un-fused logistic gradients.  R2 is R (HolE) or W (RESCAL).
    Returns dict(loss, ge, eidx, g2, idx2)."""
    n, (N, d), M = s.numel(), E.shape, R2.shape[0]
    ue, u2 = min(2 * n, N), min(n, M)
    ge, eidx, idx2, counts = _f32(ue, d), _i32(ue), _i32(u2), _i32(4)
    loss = torch.zeros(1, dtype=torch.float64, device=_ext.device())
    ws = _ws.get(lib().skge_logistic_workspace_bytes(model, n, d, N, M))
    _count('logistic_hole' if model == HOLE else 'logistic_rescal')
    if model == HOLE:
        g2 = _f32(u2, d)
        check(lib().skge_hole_logistic_grads(ptr(E), ptr(R2), ptr(s), ptr(o), ptr(p), ptr(y), ptr(valid), n, N, M, d,
                                             float(rparam), ptr(ge), ptr(eidx), ptr(g2), ptr(idx2), ptr(counts),
                                             ptr(loss), ptr(ws), ws.numel(), stream()))
    else:
        g2 = _f32(u2, d, d)
        check(lib().skge_rescal_logistic_grads(ptr(E), ptr(R2), ptr(s), ptr(o), ptr(p), ptr(y), ptr(valid), n, N, M, d,
                                               float(rparam), ptr(ge), ptr(eidx), ptr(g2), ptr(idx2), ptr(counts),
                                               ptr(loss), ptr(ws), ws.numel(), stream()))
    _, U_E, U_2, _ = counts.tolist()
    return dict(loss=float(loss.item()), ge=ge[:U_E], eidx=eidx[:U_E], g2=g2[:U_2], idx2=idx2[:U_2])


def logistic_step(model, E, R2, p2E, p2R2, s, o, p, y, rparam, opt, lr, postE, post2, counts, loss_accum,
                  ucE=None, uc2=None, valid=None):
    n, (N, d), M = s.numel(), E.shape, R2.shape[0]
    ws = _ws.get(lib().skge_logistic_workspace_bytes(model, n, d, N, M))
    fn = lib().skge_hole_logistic_step if model == HOLE else lib().skge_rescal_logistic_step
    _count('logistic_hole' if model == HOLE else 'logistic_rescal')
    check(fn(ptr(E), ptr(R2), ptr(p2E), ptr(p2R2), ptr(s), ptr(o), ptr(p), ptr(y), ptr(valid), n, N, M, d, float(rparam), opt,
             float(lr), postE, post2, ptr(counts), ptr(loss_accum), ptr(ucE), ptr(uc2), ptr(ws), ws.numel(),
             stream()))


class TripleSet(object):
    """Device hash set of training triples + corrupted-triple sampler
    (RandomModeSampler / LCWASampler of skge/sample.py)."""

    def __init__(self, s, o, p, N, M, lcwa=False):
        if N >= (1 << 24) - 1 or M > (1 << 16):
            raise ValueError('device sampler packs keys as 24/24/16 bits: need N < 2^24-1 and M <= 2^16')
        self.s, self.o, self.p, self.N, self.M = s, o, p, int(N), int(M)
        T = s.numel()
        nbytes = lib().skge_tripleset_bytes(T)
        self.table = torch.empty(nbytes, dtype=torch.uint8, device=_ext.device())
        check(lib().skge_tripleset_build(ptr(self.table), nbytes, ptr(s), ptr(o), ptr(p), T, 0, stream()))
        self.sp_table = None
        if lcwa:
            self.sp_table = torch.empty(nbytes, dtype=torch.uint8, device=_ext.device())
            check(lib().skge_tripleset_build(ptr(self.sp_table), nbytes, ptr(s), ptr(o), ptr(p), T, 1, stream()))

    def contains(self, s, o, p):
        out = torch.empty(s.numel(), dtype=torch.uint8, device=_ext.device())
        check(lib().skge_tripleset_contains(ptr(self.table), self.table.numel(), ptr(s), ptr(o), ptr(p), s.numel(),
                                            ptr(out), stream()))
        return out

    def sample(self, batch_idx, B, n_per, modes_mask, ntries, seed, offset, src=None, offset_dev=None,
               outs=None, valid=None):
        """Returns (pos, neg, valid): pos/neg are (s, o, p) int32 tensors of
        B * n_per * nmodes pairs.  ``src`` overrides the (s, o, p) arrays the
        positives are read from (default: the training arrays).  ``offset_dev`` (int64
        CUDA scalar, optional) is added to the Philox counter on the device.  ``outs`` (six
        contiguous int32 tensors or views of n entries) and ``valid`` (uint8) let the caller have the
        result written in place, e.g. into the tail of a minibatch buffer."""
        s, o, p = src if src is not None else (self.s, self.o, self.p)
        nm = bin(modes_mask & 7).count('1')
        n = B * n_per * nm
        if outs is None:
            outs = [_i32(n) for _ in range(6)]
        else:
            outs = list(outs)
            assert len(outs) == 6 and all(t.numel() == n and t.dtype == torch.int32 and t.is_contiguous() for t in outs)
        if valid is None:
            valid = torch.empty(n, dtype=torch.uint8, device=_ext.device())
        else:
            assert valid.numel() == n and valid.dtype == torch.uint8 and valid.is_contiguous()
        _count('sample')
        check(lib().skge_sample_corrupt(ptr(self.table), self.table.numel(), ptr(self.sp_table),
                                        self.sp_table.numel() if self.sp_table is not None else 0, ptr(s), ptr(o),
                                        ptr(p), ptr(batch_idx), B, n_per, modes_mask, self.N, self.M, ntries,
                                        seed & (2 ** 64 - 1), offset & (2 ** 64 - 1), ptr(offset_dev),
                                        *[ptr(t) for t in outs], ptr(valid), stream()))
        return tuple(outs[:3]), tuple(outs[3:]), valid


# ---------------------------------------------------------------------------
# ranking
# ---------------------------------------------------------------------------

def rank_op(model):
    return _ext.RANK_L1 if model == TRANSE else _ext.RANK_DOT


def make_queries(model, E, RW, kind, given, rel, target, enorm_max, coarse_rel):
    Q, d = given.numel(), E.shape[1]
    dev = _ext.device()
    q64 = torch.empty(Q, d, dtype=torch.float64, device=dev)
    q32 = torch.empty(Q, d, dtype=torch.float32, device=dev)
    tscore = torch.empty(Q, dtype=torch.float64, device=dev)
    eps, qnorm = _f32(Q), _f32(Q)
    _count('make_queries')
    check(lib().skge_rank_make_queries(model, ptr(E), ptr(RW), ptr(kind), ptr(given), ptr(rel), ptr(target), Q, d,
                                       float(enorm_max), float(coarse_rel), ptr(q64), ptr(q32), ptr(tscore),
                                       ptr(eps), ptr(qnorm), stream()))
    return dict(q64=q64, q32=q32, tscore=tscore, eps=eps, qnorm=qnorm)


def sweep_pack(X):
    """fp32 [rows, d] -> the k-major tiles of the bulk-TMA sweep (csrc/rank_sweep.cu)."""
    rows, d = X.shape
    out = torch.empty(lib().skge_rank_sweep_packed_floats(rows, d), dtype=torch.float32, device=X.device)
    _count('sweep_pack')
    check(lib().skge_rank_sweep_pack(ptr(X), rows, d, ptr(out), stream()))
    return out


def rank_sweep_tiles(op, Epk, n_shard, shard_base, d, q, Qpk, cnt_gt, cand_q, cand_e, cand_count):
    _count('sweep')
    check(lib().skge_rank_sweep_tiles(op, ptr(Epk), n_shard, shard_base, d, ptr(Qpk), ptr(q['tscore']), ptr(q['eps']),
                                      q['q32'].shape[0], ptr(cnt_gt), ptr(cand_q), ptr(cand_e), cand_q.numel(),
                                      ptr(cand_count), stream()))


def rank_rescore(op, Efull, q, pair_q, pair_e, npairs, npairs_dev, target, cnt):
    _count('rescore')
    check(lib().skge_rank_rescore(op, ptr(Efull), Efull.shape[1], ptr(q['q64']), ptr(q['tscore']), ptr(pair_q),
                                  ptr(pair_e), npairs, ptr(npairs_dev), ptr(target), ptr(cnt), stream()))


def rank_scores_one(op, E, q64_row):
    out = torch.empty(E.shape[0], dtype=torch.float64, device=_ext.device())
    check(lib().skge_rank_scores_one(op, ptr(E), E.shape[0], E.shape[1], ptr(q64_row), ptr(out), stream()))
    return out


def pack_f16(X, row_scale, scalar_scale, lo_rowmajor=False, even_tiles=False):
    """fp32 [rows, d] -> (hi, lo) fp16 UMMA blocks (uint8 tensors); with ``lo_rowmajor`` also the
    row-major copy of the lo parts (the refine-mode gather operand) and each row's squared lo norm.
    ``even_tiles`` pads ``hi`` with a zero tile to an even number of 128-row tiles (UMMA N = 256)."""
    rows, d = X.shape
    nbytes = lib().skge_rank_packed_bytes(rows, d)
    if even_tiles and ((rows + 127) // 128) % 2:
        hi = torch.zeros(lib().skge_rank_packed_bytes(rows + 128, d), dtype=torch.uint8, device=_ext.device())
    else:
        hi = torch.empty(nbytes, dtype=torch.uint8, device=_ext.device())
    lo = torch.empty(nbytes, dtype=torch.uint8, device=_ext.device())
    lo_rm = torch.empty(nbytes, dtype=torch.uint8, device=_ext.device()) if lo_rowmajor else None
    lo_n2 = torch.zeros(rows, dtype=torch.float32, device=_ext.device()) if lo_rowmajor else None
    _count('pack')
    check(lib().skge_rank_pack_f16(ptr(X), rows, d, ptr(row_scale), float(scalar_scale), ptr(hi), ptr(lo),
                                   ptr(lo_rm), ptr(lo_n2), stream()))
    return (hi, lo, lo_rm, lo_n2) if lo_rowmajor else (hi, lo)


def query_scale(q, escale):
    Q, d = q['q32'].shape
    qscale, tlo, thi = _f32(Q), _f32(Q), _f32(Q)
    _count('pack')
    check(lib().skge_rank_query_scale(ptr(q['q32']), ptr(q['tscore']), ptr(q['eps']), Q, d, float(escale),
                                      ptr(qscale), ptr(tlo), ptr(thi), stream()))
    return qscale, tlo, thi


def rank_gemm_count(Ehi, Elo, n_shard, shard_base, Qhi, Qlo, Q, d, nsplit, tlo, thi, cnt_gt, cand_q, cand_e,
                    cand_count, qwidth=None, tile_w=None, perm=None, lo_scale=None, q1w=None):
    _count('gemm')
    check(lib().skge_rank_gemm_count(ptr(Ehi), ptr(Elo), n_shard, shard_base, ptr(Qhi), ptr(Qlo), Q, d, nsplit,
                                     ptr(tlo), ptr(thi), ptr(qwidth), ptr(tile_w), ptr(perm), ptr(lo_scale),
                                     ptr(q1w), ptr(cnt_gt), ptr(cand_q), ptr(cand_e), cand_q.numel(),
                                     ptr(cand_count), stream()))


def quant_lo(lo_rm, rows, d):
    """Row-major fp16 lo rows -> (uint8 rows, fp32 scale per padded row)."""
    rp = (rows + 127) // 128 * 128
    kb = (d + 63) // 64 * 64
    lo8 = torch.empty(rp * kb, dtype=torch.uint8, device=_ext.device())
    scale = torch.empty(rp, dtype=torch.float32, device=_ext.device())
    _count('pack')
    check(lib().skge_rank_quant_lo(ptr(lo_rm), rows, d, ptr(lo8), ptr(scale), stream()))
    return lo8, scale


def quant_lo_s8(lo_rm, rows, d):
    """Row-major fp16 lo rows -> (int8 rows, float32 [padded rows, 2] = (scale, ||lo||_1))."""
    rp = (rows + 127) // 128 * 128
    kb = (d + 63) // 64 * 64
    lo8 = torch.empty(rp * kb, dtype=torch.int8, device=_ext.device())
    meta = torch.empty(rp, 2, dtype=torch.float32, device=_ext.device())
    _count('pack')
    check(lib().skge_rank_quant_lo_s8(ptr(lo_rm), rows, d, ptr(lo8), ptr(meta), stream()))
    return lo8, meta


def pack_q8(q, qscale, tlo, thi):
    """Swizzled int8 query tiles + the per-query constants of the refine epilogue."""
    Q, d = q['q32'].shape
    kb = (d + 63) // 64 * 64
    qtiles = (Q + 127) // 128
    Q8 = torch.empty(qtiles * 128 * kb, dtype=torch.int8, device=_ext.device())
    qmeta = torch.empty(Q, 8, dtype=torch.float32, device=_ext.device())
    _count('pack')
    check(lib().skge_rank_pack_q8(ptr(q['q32']), ptr(qscale), ptr(q['qnorm']), ptr(tlo), ptr(thi), Q, d, ptr(Q8),
                                  ptr(qmeta), stream()))
    return Q8, qmeta


def rank_refine_count(Ehi, Elo8, lo_meta, tile_w, perm, n_shard, shard_base, Qhi, Qlo, Q8, qmeta, Q, d, cta_group,
                      cnt_gt, cand_q, cand_e, cand_count):
    _count('gemm')
    check(lib().skge_rank_refine_count(ptr(Ehi), ptr(Elo8), ptr(lo_meta), ptr(tile_w), ptr(perm), n_shard,
                                       shard_base, ptr(Qhi), ptr(Qlo), ptr(Q8), ptr(qmeta), Q, d, cta_group,
                                       ptr(cnt_gt), ptr(cand_q), ptr(cand_e), cand_q.numel(), ptr(cand_count),
                                       stream()))


def quant_rows(X, scale):
    """fp32 rows (packed order) -> (E8 int8 [padded rows, 2, kb], meta float32 [padded rows, 4],
    norms float32 [padded rows, 2]) for the single-product engine; rows padded to a multiple of 256."""
    rows, d = X.shape
    rp = (rows + 255) // 256 * 256
    kb = (d + 63) // 64 * 64
    E8 = torch.empty(rp, 2, kb, dtype=torch.int8, device=_ext.device())
    meta = torch.empty(rp, 4, dtype=torch.float32, device=_ext.device())
    norms = torch.empty(rp, 2, dtype=torch.float32, device=_ext.device())
    _count('pack')
    check(lib().skge_rank_quant_rows(ptr(X), rows, d, float(scale), ptr(E8), ptr(meta), ptr(norms), stream()))
    return E8, meta, norms


def pack_q8x2(q, qscale, tlo, thi):
    """Swizzled int8 tiles of the queries' hi and lo parts + the per-query constants of the
    single-product epilogue (16 floats per query, allocated for whole 128-query tiles)."""
    Q, d = q['q32'].shape
    kb = (d + 63) // 64 * 64
    qtiles = (Q + 127) // 128
    Q8h = torch.empty(qtiles * 128 * kb, dtype=torch.int8, device=_ext.device())
    Q8l = torch.empty(qtiles * 128 * kb, dtype=torch.int8, device=_ext.device())
    qmeta = torch.zeros(qtiles * 128, 16, dtype=torch.float32, device=_ext.device())
    _count('pack')
    check(lib().skge_rank_pack_q8x2(ptr(q['q32']), ptr(qscale), ptr(tlo), ptr(thi), Q, d, ptr(Q8h), ptr(Q8l),
                                    ptr(qmeta), stream()))
    return Q8h, Q8l, qmeta


def rank_single_count(Ehi, E8, e_meta, tile_w, perm, n_shard, shard_base, Qhi, Q8h, Q8l, qmeta, Q, d, cta_group,
                      cnt_gt, cand_q, cand_e, cand_count):
    _count('gemm')
    check(lib().skge_rank_single_count(ptr(Ehi), ptr(E8), ptr(e_meta), ptr(tile_w), ptr(perm), n_shard, shard_base,
                                       ptr(Qhi), ptr(Q8h), ptr(Q8l), ptr(qmeta), Q, d, cta_group, ptr(cnt_gt),
                                       ptr(cand_q), ptr(cand_e), cand_q.numel(), ptr(cand_count), stream()))
