"""Filtered link-prediction ranking on the device.

Reference: FilteredRankingEval (skge/base.py:739-759, 913-1031), the per-model
scorers TransEEval (skge/run_transe.py:13-29) and HolEEval
(skge/run_hole.py:10-19), ranking_scores / compute_scores
(skge/base.py:1050-1103).

The reference scores one query at a time against all N entities and argsorts
four times per test triple.  Here every query of the test set is one row of a
query matrix; a coarse sweep over the (sharded) entity table counts, per query,
the entities that beat the target by more than an error bound and lists the
undecided ones; those and the filter entries are settled in fp64 from the fp32
master table.  rank = 1 + #{score > target score}, which equals the reference's
argsort position whenever the target's score is not tied.

With torch.distributed initialised (one process per GPU) the entity table is
row-partitioned across ranks, each rank counts over its own shard and one
all-reduce of 2Q int32 counts combines them; results are identical on every
rank and for every world size.
"""
import logging
import os
import math

import numpy as np
import torch

from . import _ext, kernels

log = logging.getLogger('EX-KG')


# ---------------------------------------------------------------------------
# host-side logic shared by all world sizes (pure functions: CPU-testable)
# ---------------------------------------------------------------------------

def shard_range(N, rank, world):
    """Row range [lo, hi) of the entity table owned by ``rank``: contiguous
    blocks of ceil(N / world) rows."""
    per = (N + world - 1) // world
    lo = min(N, rank * per)
    return lo, min(N, lo + per)


def flatten_queries(test):
    """(Te, 3) test triples (s, o, p) -> query arrays of length 2*Te:
    first the Te tail queries (s, p, ?) -> o, then the Te head queries (?, p, o) -> s.
    Returns kind (uint8), given, rel, target (int64 numpy)."""
    t = np.asarray(test, dtype=np.int64).reshape(-1, 3)
    te = t.shape[0]
    kind = np.concatenate([np.zeros(te, np.uint8), np.ones(te, np.uint8)])
    given = np.concatenate([t[:, 0], t[:, 1]])
    rel = np.concatenate([t[:, 2], t[:, 2]])
    target = np.concatenate([t[:, 1], t[:, 0]])
    return kind, given, rel, target


def build_filter_pairs(true_triples, kind, given, rel, target, device=None):
    """All (query, entity) filter entries: for a tail query (s, p, ?) the known
    objects of (s, p) other than the target, for a head query the known subjects
    of (p, o) (skge/base.py:744-752, 970-977, 1012-1014).  Duplicates collapse
    (the reference writes -inf twice).  Returns int32 tensors (pair_q ascending,
    pair_e) on ``device``; torch ops only, so it runs on CPU or GPU."""
    dev = device if device is not None else torch.device('cpu')
    if len(given) == 0:
        z = torch.zeros(0, dtype=torch.int32, device=dev)
        return z, z.clone()
    if isinstance(true_triples, torch.Tensor):
        tt = true_triples.to(device=dev, dtype=torch.int64).reshape(-1, 3)
    else:
        tt = torch.as_tensor(np.asarray(true_triples, dtype=np.int64).reshape(-1, 3), device=dev)
    kind_t = torch.as_tensor(kind.astype(np.int64), device=dev)
    given_t = torch.as_tensor(given, device=dev)
    rel_t = torch.as_tensor(rel, device=dev)
    target_t = torch.as_tensor(target, device=dev)
    nmax = int(max(tt[:, :2].max().item() if tt.numel() else 0, given_t.max().item(), target_t.max().item())) + 1
    out_q, out_e = [], []
    for k, (gcol, vcol) in enumerate(((0, 1), (1, 0))):   # tail: key (p, s) -> o ; head: key (p, o) -> s
        qsel = torch.nonzero(kind_t == k).flatten()
        if qsel.numel() == 0 or tt.numel() == 0:
            continue
        keys = tt[:, 2] * nmax + tt[:, gcol]
        if (int(tt[:, 2].max().item()) + 1) * nmax * nmax < 2 ** 62:
            packed = torch.unique(keys * nmax + tt[:, vcol])     # sorted, duplicates collapsed
            ukeys, uvals = packed // nmax, packed % nmax
        else:
            # (p, given, value) does not fit one int64 key: sort (key, value) rows lexicographically
            rows = torch.unique(torch.stack([keys, tt[:, vcol]], 1), dim=0)
            ukeys, uvals = rows[:, 0].contiguous(), rows[:, 1].contiguous()
        qkey = rel_t[qsel] * nmax + given_t[qsel]
        lo = torch.searchsorted(ukeys, qkey, right=False)
        hi = torch.searchsorted(ukeys, qkey, right=True)
        lens = hi - lo
        total = int(lens.sum().item())
        if total == 0:
            continue
        owner = torch.repeat_interleave(torch.arange(qsel.numel(), device=dev), lens)
        start = torch.cumsum(lens, 0) - lens
        pos = lo[owner] + (torch.arange(total, device=dev) - start[owner])
        e = uvals[pos]
        q = qsel[owner]
        keep = e != target_t[q]
        out_q.append(q[keep])
        out_e.append(e[keep])
    if not out_q:
        z = torch.zeros(0, dtype=torch.int32, device=dev)
        return z, z.clone()
    q = torch.cat(out_q)
    e = torch.cat(out_e)
    order = torch.argsort(q, stable=True)
    return q[order].to(torch.int32), e[order].to(torch.int32)


def allreduce_counts(cnt):
    """Sum the per-shard counts over ranks (no-op without a process group)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    return cnt


def ranks_from_counts(cnt):
    """cnt: int32 [2, Q] = (#entities beating the target, #filter entities
    beating the target) -> (raw ranks, filtered ranks)."""
    raw = 1 + cnt[0]
    return raw, raw - cnt[1]


def regroup_plan(test):
    """The part of ``regroup`` that depends on the test triples only (computed once per
    evaluator): the permutation that groups triples by relation, the group boundaries, and the
    relations in order of first appearance."""
    t = np.asarray(test, dtype=np.int64).reshape(-1, 3)
    te = t.shape[0]
    if te == 0:
        return None
    rel = t[:, 2]
    if rel.max() < 65536:
        rel = rel.astype(np.uint16)                     # numpy's stable sort is a radix sort for 16-bit keys
    order = np.argsort(rel, kind='stable')              # triples of one relation stay in test order
    ps = t[order, 2]
    bounds = np.concatenate([[0], np.nonzero(np.diff(ps))[0] + 1, [te]]).tolist()
    first = order[bounds[:-1]]                          # first test position of each relation
    groups = [(int(ps[bounds[g]]), bounds[g], bounds[g + 1]) for g in np.argsort(first, kind='stable').tolist()]
    return te, np.concatenate([order, te + order]), groups


class _RelRanks(dict):
    """``{'head': [...], 'tail': [...]}`` of one relation whose Python lists are built from the
    numpy rank arrays on first use: a plain dict for every reader (equality, iteration, pickling),
    but a ranking pass over 1k relations does not pay for 400k int objects nobody looks at."""

    __slots__ = ('_h', '_t')

    def __init__(self, head, tail):
        dict.__init__(self)
        self._h, self._t = head, tail

    def _fill(self):
        if self._h is not None:
            h, t = self._h, self._t
            self._h = self._t = None
            dict.__setitem__(self, 'head', h.tolist())
            dict.__setitem__(self, 'tail', t.tolist())
        return self

    def __missing__(self, key):
        if self._h is None:
            raise KeyError(key)
        return dict.__getitem__(self._fill(), key)

    def __getitem__(self, key):
        return dict.__getitem__(self._fill(), key)

    def get(self, key, default=None):
        return dict.get(self._fill(), key, default)

    def keys(self):
        return dict.keys(self._fill())

    def values(self):
        return dict.values(self._fill())

    def items(self):
        return dict.items(self._fill())

    def __iter__(self):
        return dict.__iter__(self._fill())

    def __len__(self):
        return dict.__len__(self._fill())

    def __contains__(self, key):
        return dict.__contains__(self._fill(), key)

    def __eq__(self, other):
        if isinstance(other, _RelRanks):
            other._fill()
        return dict.__eq__(self._fill(), other)

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = None

    def __repr__(self):
        return dict.__repr__(self._fill())

    def __setitem__(self, key, value):
        dict.__setitem__(self._fill(), key, value)

    def __delitem__(self, key):
        dict.__delitem__(self._fill(), key)

    def pop(self, *a):
        return dict.pop(self._fill(), *a)

    def setdefault(self, *a):
        return dict.setdefault(self._fill(), *a)

    def update(self, *a, **k):
        dict.update(self._fill(), *a, **k)

    def copy(self):
        return dict(self._fill())

    def __reduce__(self):
        return (dict, (dict(self._fill()),))


def regroup(test, raw, filt, plan=None):
    """Flat rank arrays (tail queries then head queries) -> the reference's
    ``pos`` / ``fpos`` dicts ``{p: {'head': [...], 'tail': [...]}}`` with
    relations and triples in insertion order (skge/base.py:743, 921, 1027-1028).
    The per-relation lists are materialised on first access (``_RelRanks``)."""
    plan = plan or regroup_plan(test)
    pos, fpos = {}, {}
    if plan is None:
        return pos, fpos
    te, order2, groups = plan
    # one permutation per array; a relation's ranks are then contiguous slices
    r, f = np.asarray(raw)[order2], np.asarray(filt)[order2]
    for p, a, b in groups:                                  # relations in order of first appearance
        pos[p] = _RelRanks(r[te + a:te + b], r[a:b])
        fpos[p] = _RelRanks(f[te + a:te + b], f[a:b])
    return pos, fpos


def _world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


# ---------------------------------------------------------------------------
# evaluators
# ---------------------------------------------------------------------------

class FilteredRankingEval(object):
    """FilteredRankingEval(xs, true_triples, neval=-1).positions(mdl) -> (pos, fpos).

    Subclasses name the model family (``model_code``); the reference's
    per-query hooks ``prepare`` / ``scores_o`` / ``scores_s`` are still offered
    (computed on the device in fp64) for callers that want raw score vectors.
    """

    model_code = None
    chunk_queries = 1 << 18     # queries per device pass (bounds q64/q32 scratch: 3 KB per query at d = 256)
    use_tensor_cores = True     # DOT models: tcgen05 coarse pass when the shapes allow

    def __init__(self, xs, true_triples, neval=-1):
        if isinstance(xs, torch.Tensor):
            xs = xs.cpu().numpy()
        self.test = np.asarray(xs, dtype=np.int64).reshape(-1, 3)
        self._true = true_triples
        self.sz = len(self.test)
        self.neval = neval
        self.kind, self.given, self.rel, self.target = flatten_queries(self.test)
        self._dev = None
        self.last_stats = {}

    # the reference's python indices, built on demand (small graphs only)
    @property
    def idx(self):
        d = {}
        for s, o, p in self.test.tolist():
            d.setdefault(p, []).append((s, o))
        return d

    @property
    def tt(self):
        d = {}
        for s, o, p in np.asarray(self._true, dtype=np.int64).reshape(-1, 3).tolist():
            e = d.setdefault(p, {'ss': {}, 'os': {}})
            e['os'].setdefault(s, []).append(o)
            e['ss'].setdefault(o, []).append(s)
        return d

    # -- device state, built once ------------------------------------------------
    def _device_state(self):
        """The filter index (built once, on the device) plus this call's upload
        of the query descriptors from pinned host memory."""
        dev = _ext.device()
        if self._dev is None:
            pq, pe = build_filter_pairs(self._true, self.kind, self.given, self.rel, self.target, device=dev)
            host = dict(kind=torch.from_numpy(self.kind),
                        given=torch.from_numpy(self.given.astype(np.int32)),
                        rel=torch.from_numpy(self.rel.astype(np.int32)),
                        target=torch.from_numpy(self.target.astype(np.int32)))
            self._host = {k: v.pin_memory() for k, v in host.items()}
            self._dev = dict(pair_q=pq.contiguous(), pair_e=pe.contiguous())
        for k, v in self._host.items():
            self._dev[k] = v.to(dev, non_blocking=True)
        return self._dev

    def _check_ids(self, N, M):
        """Out-of-range ids are an IndexError in the reference (fancy indexing, skge/base.py:944-966);
        the kernels do not bounds-check, so the test and filter triples are validated once."""
        if getattr(self, '_ids_ok', None) == (N, M):
            return
        for name, a in (('test', self.test), ('true', self._true)):
            if isinstance(a, torch.Tensor):
                if a.numel() == 0:
                    continue
                a2 = a.reshape(-1, 3)
                lo, hi_e, hi_p = int(a2.min()), int(a2[:, :2].max()), int(a2[:, 2].max())
            else:
                a2 = np.asarray(a, dtype=np.int64).reshape(-1, 3)
                if a2.size == 0:
                    continue
                lo, hi_e, hi_p = int(a2.min()), int(a2[:, :2].max()), int(a2[:, 2].max())
            if lo < 0 or hi_e >= N or hi_p >= M:
                raise IndexError('%s triples: ids out of range for %d entities / %d relations' % (name, N, M))
        self._ids_ok = (N, M)

    def h2d_bytes(self):
        return int(sum(v.numel() * v.element_size() for v in self._host.values())) if self._dev else 0

    def _second(self, mdl):
        return mdl.W.data if self.model_code == _ext.MODEL_RESCAL else mdl.R.data

    # -- the ranking pass ------------------------------------------------------------
    def positions(self, mdl, plot=False, pagerankMap=None):
        if self.model_code is None:
            raise NotImplementedError('derive from TransEEval / HolEEval / RESCALEval: the device ranking '
                                      'pass needs to know the model family (there is no host fallback)')
        if self.sz == 0:
            return {}, {}
        cnt = self.count_pass(mdl).cpu().numpy()           # one device -> host copy of [2, Q] int32
        raw = 1 + cnt[0]
        if getattr(self, '_regroup_plan', None) is None:
            self._regroup_plan = regroup_plan(self.test)
        return regroup(self.test, raw, raw - cnt[1], self._regroup_plan)

    def count_pass(self, mdl, E=None, world=None):
        """int32 [2, Q] counts, already summed over ranks.  ``world`` = (rank,
        world_size) overrides the process group and skips the reduction (used to
        emulate several shards on one GPU).

        The host synchronises twice per pass: once for the table's checksum (which validates the
        cached fp16 / int8 shadow of the shard and the cached row-norm bound) and once at the end
        for the candidate counts (list overflow -> grow and redo the pass)."""
        st = self._device_state()
        dev = _ext.device()
        E = mdl.E.data if E is None else E
        RW = self._second(mdl)
        N, d = E.shape
        Q = st['given'].numel()
        self._check_ids(N, RW.shape[0])
        emulated = world is not None
        rank, world = world if emulated else _world()
        lo, hi = shard_range(N, rank, world)
        op = kernels.rank_op(self.model_code)
        _PASS.clear()
        stats = _table_stats(E)           # the pass's one checksum read of the table (and its host sync)
        _PASS.update(key=(E.data_ptr(), tuple(E.shape)), stats=stats)
        enorm = stats['enorm'] if op == _ext.RANK_DOT else 1.0
        # filter entries are settled by the rank that owns the entity
        pq_chunks, pe_chunks, npairs = self._shard_pairs(st, lo, hi, world, Q)
        engine = self._coarse_engine(E, lo, hi, enorm, min(Q, self.chunk_queries))
        engine.reserve(min(Q, self.chunk_queries))
        while True:
            cnt = torch.zeros(2, Q, dtype=torch.int32, device=dev)
            engine.begin_pass()
            for ci, q0 in enumerate(range(0, Q, self.chunk_queries)):
                q1 = min(Q, q0 + self.chunk_queries)
                sl = slice(q0, q1)
                q = kernels.make_queries(self.model_code, E, RW, st['kind'][sl], st['given'][sl], st['rel'][sl],
                                         st['target'][sl], enorm, engine.coarse_rel(d))
                engine.run(op, q, cnt[0, sl])
                if pq_chunks[ci].numel():
                    kernels.rank_rescore(op, E, q, pq_chunks[ci], pe_chunks[ci], pq_chunks[ci].numel(), None, None,
                                         cnt[1, sl])
            ncand, worst = engine.end_pass()
            if worst <= engine.cap:
                break
            engine.grow(worst + 1)       # a candidate list overflowed: redo the pass with a larger one
        _PASS.clear()
        self.last_stats = dict(candidates=ncand, filter_pairs=npairs, shard=(lo, hi), world=world,
                               engine=engine.name, dtype=engine.dtype)
        return cnt if emulated else allreduce_counts(cnt)

    def _shard_pairs(self, st, lo, hi, world, Q):
        """This shard's filter pairs, cut per query chunk with chunk-local query ids (the index does
        not change between passes, so this is done once per shard)."""
        key = (lo, hi, world, self.chunk_queries)
        cache = self.__dict__.setdefault('_pair_cache', {})
        hit = cache.get(key)
        if hit is None:
            pq, pe = st['pair_q'], st['pair_e']
            if world > 1:
                own = (pe >= lo) & (pe < hi)
                pq, pe = pq[own].contiguous(), pe[own].contiguous()
            edges = torch.arange(0, Q + self.chunk_queries, self.chunk_queries, device=pq.device)
            bounds = torch.searchsorted(pq.to(torch.int64), edges).tolist()
            pqs = [(pq[a:b] - q0).contiguous()
                   for q0, a, b in zip(range(0, Q, self.chunk_queries), bounds[:-1], bounds[1:])]
            pes = [pe[a:b].contiguous() for a, b in zip(bounds[:-1], bounds[1:])]
            if len(cache) >= 16:
                cache.clear()
            hit = cache[key] = (pqs, pes, int(pq.numel()))
        return hit

    # 'auto' | 'sweep' (fp32 CUDA cores) | 'umma' (tcgen05, DOT models, d <= 256; nsplit products) |
    # 'single' (tcgen05, one product + int8 refinement of both cross terms)
    engine = 'auto'
    # fp16 hi/lo products on the tensor cores: 3, 2 (third product added in the epilogue), 1 (fp16 only);
    # 0 = choose: 2 for large sweeps (the saved MMA work outweighs the heavier epilogue), else 3
    nsplit = int(os.environ.get('SKGE_RANK_NSPLIT', '0'))
    # nsplit = 2 only: 2 pairs two CTAs on one 256-query x 256-entity MMA (cta_group::2), 1 = one CTA per MMA
    cta_group = int(os.environ.get('SKGE_RANK_CG', '2'))
    refine_min_pairs = 1 << 28  # queries x shard rows per coarse launch above which nsplit = 2 pays off (config 2: 2^28.6)
    # what 'auto' runs above that size: 'refine' (two products + one refined cross term, csrc/rank_refine.cu)
    # or 'single' (one product + both cross terms refined, csrc/rank_single.cu)
    large_sweep_engine = os.environ.get('SKGE_RANK_LARGE', 'refine')

    def _coarse_engine(self, E, lo, hi, enorm, nqueries=0):
        """The coarse-pass engine for this shard.  The object (and its candidate
        buffers) is cached across calls; the fp16 shadow of the shard is rebuilt on
        every call because the model may have been trained in between."""
        dot = kernels.rank_op(self.model_code) == _ext.RANK_DOT
        want = self.engine
        if want == 'auto':
            want = 'umma' if (dot and self.use_tensor_cores and E.shape[1] <= 256) else 'sweep'
        if want in ('umma', 'single') and (not dot or E.shape[1] > 256):
            raise ValueError('the tcgen05 engines need a dot-product model with d <= 256')
        nsplit = self.nsplit
        if not nsplit:
            nsplit = 3
            if want == 'umma' and nqueries * (hi - lo) >= self.refine_min_pairs:
                nsplit = 2
                if self.large_sweep_engine == 'single':
                    want = 'single'
        if want == 'single':
            nsplit = 1
        key = (want, nsplit if want != 'sweep' else 0, self.cta_group if want == 'single' or nsplit == 2 else 0)
        cache = self.__dict__.setdefault('_engines', {})
        eng = cache.get(key)
        if eng is None:
            eng = cache[key] = (_SweepEngine() if want == 'sweep' else
                                _UmmaEngine(nsplit, self.cta_group, single=(want == 'single')))
        eng.bind(E, lo, hi)
        return eng

    # -- reference-style hooks (fp64 score vectors from the device) ---------------------
    def prepare(self, mdl, p):
        self._hook_p = p

    def _scores_one(self, mdl, given, p, kind):
        dev = _ext.device()
        one = lambda v, dt: torch.tensor([v], dtype=dt, device=dev)  # noqa: E731
        q = kernels.make_queries(self.model_code, mdl.E.data, self._second(mdl), one(kind, torch.uint8),
                                 one(given, torch.int32), one(p, torch.int32), one(0, torch.int32), 1.0, 0.0)
        return kernels.rank_scores_one(kernels.rank_op(self.model_code), mdl.E.data, q['q64'][0]).cpu().numpy()

    def scores_o(self, mdl, s, p):
        return self._scores_one(mdl, s, p, 0)

    def scores_s(self, mdl, o, p):
        return self._scores_one(mdl, o, p, 1)


TIMINGS = []   # (start event, end event, algorithmic flops/ops) per coarse launch when timing is on

_STATS = {}


def _checksum(t):
    """Wrapping 64-bit sum of a float32 tensor's bit patterns: one read of the data, one host sync.
    Summed as int64 words where the layout allows (no widening copy of the table)."""
    t = t.detach()
    if t.is_contiguous() and t.numel() % 2 == 0 and t.data_ptr() % 8 == 0:
        return int(t.view(-1).view(torch.int64).sum().item())
    return int(torch.sum(t.reshape(-1).view(torch.int32), dtype=torch.int64).item())


def _table_stats(E):
    """Checksum (wrapping int64 sum of the fp32 bit patterns) and largest row norm of a table.
    The checksum costs one read of the table and one host sync per ranking pass; it validates
    everything that is derived from the parameters and cached across passes (row-norm bound, the
    shard's fp16 / int8 shadow), whoever changed the table and however."""
    chk = _checksum(E)
    key = (E.data_ptr(), tuple(E.shape))
    hit = _STATS.get(key)
    if hit is None or hit['chk'] != chk:
        if len(_STATS) >= 8:
            _STATS.clear()
        hit = _STATS[key] = dict(chk=chk, enorm=float(torch.linalg.vector_norm(E, dim=1).max().item()))
    return hit


_PASS = {}


def _pass_stats(E):
    """_table_stats computed once per ranking pass: count_pass() stores it here before binding the
    engine, so the engine's shadow validation does not read the table a second time."""
    hit = _PASS.get('stats')
    if hit is not None and _PASS.get('key') == (E.data_ptr(), tuple(E.shape)):
        return hit
    return _table_stats(E)


class _SweepEngine(object):
    """fp32 coarse sweep on the CUDA cores (any model, any d) + fp64 settlement."""
    name = 'fp32-sweep'
    dtype = 'f32 sweep + f64 settle'
    cands_per_query = 512       # initial sizing of the candidate list

    def __init__(self):
        self.cap = 0
        self.cand_q = self.cand_e = self.count = None
        self._counts = []

    def bind(self, E, lo, hi):
        self.E = E
        self.lo, self.hi = lo, hi
        self.shard = E[lo:hi]
        if self.count is None or self.count.device != E.device:
            self.count = torch.zeros(1, dtype=torch.int64, device=E.device)
            self.cap = 0
        if type(self) is _SweepEngine and hi > lo:
            # the k-major packed copy (csrc/rank_sweep.cu) of the shard is rebuilt only when the table's checksum changes
            # the whole table's checksum (already computed for this pass) also validates a shard's shadow
            chk = _pass_stats(E)['chk']
            key = (E.data_ptr(), lo, hi, E.shape[1], chk)
            if key != getattr(self, '_pack_key', None):
                self._pack_key = None
                self.Epk = kernels.sweep_pack(self.shard)
                self._pack_key = key

    def reserve(self, nqueries):
        self.grow(max(1 << 20, self.cands_per_query * nqueries))

    def grow(self, n):
        if n > self.cap:
            self.cap = 1 << int(math.ceil(math.log2(n)))
            self.cand_q = torch.empty(self.cap, dtype=torch.int32, device=self.E.device)
            self.cand_e = torch.empty(self.cap, dtype=torch.int32, device=self.E.device)

    def coarse_rel(self, d):
        return 2.0 * (d + 2) * 2.0 ** -24

    # CUDA-event timing of the coarse kernel alone (bench.py's roofline leg)
    timing = False

    def _timed(self, fn, work):
        if not self.timing:
            return fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        TIMINGS.append((a, b, work))

    def _coarse(self, op, q, cnt_gt):
        d = self.E.shape[1]
        work = 2.0 * (self.hi - self.lo) * d * q['q32'].shape[0]
        Qpk = kernels.sweep_pack(q['q32'])
        self._timed(lambda: kernels.rank_sweep_tiles(op, self.Epk, self.hi - self.lo, self.lo, d, q, Qpk, cnt_gt,
                                                     self.cand_q, self.cand_e, self.count), work)

    def begin_pass(self):
        self._counts = []

    def run(self, op, q, cnt_gt):
        """Adds this shard's counts for the query chunk into cnt_gt.  Nothing here waits for the
        device: the settlement reads the candidate count from device memory and the host looks at
        it once per pass (end_pass)."""
        if self.hi <= self.lo:
            return
        self.count.zero_()
        self._coarse(op, q, cnt_gt)
        kernels.rank_rescore(op, self.E, q, self.cand_q, self.cand_e, self.cap, self.count, None, cnt_gt)
        self._counts.append(self.count.clone())

    def end_pass(self):
        """(band candidates settled in fp64 over the pass, largest per-chunk count)."""
        if not self._counts:
            return 0, 0
        c = torch.cat(self._counts).tolist()
        self._counts = []
        return int(sum(c)), int(max(c))


class _UmmaEngine(_SweepEngine):
    """tcgen05 coarse pass: fp16 hi/lo split operands, fp32 accumulation in TMEM,
    count / band epilogue straight out of TMEM.  nsplit = 3 / 1: csrc/rank_umma.cu; nsplit = 2
    (large sweeps): csrc/rank_refine.cu, two products on the tensor cores and the third one added
    from int8 copies of both operands for the pairs near the boundary.  The shadow of the shard is
    rebuilt only when the table's checksum changes."""
    name = 'tcgen05-f16x3'
    cands_per_query = 160

    def __init__(self, nsplit=3, cta_group=1, single=False):
        super(_UmmaEngine, self).__init__()
        self.nsplit = nsplit
        self.cta_group = cta_group
        self.single = single
        self.name = 'tcgen05-f16x%d' % nsplit
        self.dtype = 'f16x%d split (tcgen05, fp32 accumulate) + f64 settle' % nsplit
        if single:
            self.name = 'tcgen05-f16x1-refined'
            self.dtype = 'f16 (tcgen05, fp32 accumulate) + int8 dp4a refinement of both cross terms + f64 settle'
            self.cands_per_query = 320
            self._shadow_key = None
            return
        if nsplit == 2:
            self.dtype = 'f16x2 split (tcgen05, fp32 accumulate) + int8 dp4a refinement + f64 settle'
            self.cands_per_query = 256
        if nsplit == 1:
            self.cands_per_query = 4096
        self._shadow_key = None

    def bind(self, E, lo, hi):
        super(_UmmaEngine, self).bind(E, lo, hi)
        if hi <= lo:
            return
        # the whole table's checksum (already computed for this pass) also validates a shard's shadow
        chk = _pass_stats(E)['chk']
        key = (E.data_ptr(), lo, hi, E.shape[1], chk)
        if key == self._shadow_key:
            return
        self._shadow_key = None
        emax = float(self.shard.abs().max().item())
        self.escale = 2.0 ** (12 - math.ceil(math.log2(emax))) if emax > 0 else 1.0
        if self.single:
            # packed by decreasing row norm like nsplit = 2 (see below); int8 copies of the lo AND hi parts
            rn = torch.linalg.vector_norm(self.shard, dim=1)
            self.perm = torch.argsort(rn, descending=True).to(torch.int32)
            ordered = self.shard.index_select(0, self.perm.to(torch.int64))
            self.Ehi, _ = kernels.pack_f16(ordered, None, self.escale, even_tiles=True)
            self.E8, self.e_meta, norms = kernels.quant_rows(ordered, self.escale)
            del ordered, _
            nl, nh = norms[:, 0].reshape(-1, 128), norms[:, 1].reshape(-1, 128)
            wl = nl.max(dim=1).values * 1.001
            wh = (nh + nl).max(dim=1).values * 1.001
            self.tile_w = torch.stack([wl, wh], 1).contiguous()
        elif self.nsplit == 2:
            # The epilogue gathers lo rows and widens the band of a 128-row entity tile by
            # ||q|| * max ||e_lo|| over the tile.  Counting does not care about the order of the
            # entities, so the shard is packed by decreasing row norm: the rows of a tile are then
            # alike and the per-tile bound is tight even when the table mixes long and short rows.
            rn = torch.linalg.vector_norm(self.shard, dim=1)
            self.perm = torch.argsort(rn, descending=True).to(torch.int32)
            ordered = self.shard.index_select(0, self.perm.to(torch.int64))
            self.Ehi, _, lo_rm, n2 = kernels.pack_f16(ordered, None, self.escale, lo_rowmajor=True, even_tiles=True)
            self.Elo8, self.lo_meta = kernels.quant_lo_s8(lo_rm, ordered.shape[0], ordered.shape[1])
            del ordered, lo_rm
            pad = (-n2.numel()) % 256
            if pad:
                n2 = torch.cat([n2, n2.new_zeros(pad)])
            # 1e-3: fp32 atomics' rounding in the squared norms
            self.tile_w = (n2.view(-1, 128).max(dim=1).values.sqrt() * 1.001).contiguous()
        else:
            self.Ehi, self.Elo = kernels.pack_f16(self.shard, None, self.escale)
        self._shadow_key = key

    def coarse_rel(self, d):
        # Error model, relative to sum|q_i e_i| <= |q||e|: split residual 3*2^-22, fp32 rounding
        # of q 2^-24, and at most 2^-23 per tcgen05.mma over the 3*d/16 <= 48 accumulation steps
        # (worst case, all one-sided): 6.2e-6 at d = 256.  2^-17 = 7.6e-6 covers it; the error
        # measured on the B200 (profiles/exp_gemm.py probe) is 1.1e-7 = 2^-23.1, 70x smaller.
        # nsplit = 1 keeps only hi*hi: fp16 rounding of both operands, 2^-10 worst case.
        # nsplit = 2 adds the third product in the epilogue from int8 copies of both operands; the
        # quantisation error of that term is bounded per pair and added to that pair's band there.
        return 2.0 ** -9 if (self.nsplit == 1 and not self.single) else 2.0 ** -17

    def _coarse(self, op, q, cnt_gt):
        Q, d = q['q32'].shape
        qscale, tlo, thi = kernels.query_scale(q, self.escale)
        Qhi, Qlo = kernels.pack_f16(q['q32'], qscale, 1.0)
        work = 2.0 * (self.hi - self.lo) * d * Q
        if self.single:
            Q8h, Q8l, qmeta = kernels.pack_q8x2(q, qscale, tlo, thi)
            self._timed(lambda: kernels.rank_single_count(self.Ehi, self.E8, self.e_meta, self.tile_w, self.perm,
                                                          self.hi - self.lo, self.lo, Qhi, Q8h, Q8l, qmeta, Q, d,
                                                          self.cta_group, cnt_gt, self.cand_q, self.cand_e,
                                                          self.count), work)
        elif self.nsplit == 2:
            Q8, qmeta = kernels.pack_q8(q, qscale, tlo, thi)
            self._timed(lambda: kernels.rank_refine_count(self.Ehi, self.Elo8, self.lo_meta, self.tile_w, self.perm,
                                                          self.hi - self.lo, self.lo, Qhi, Qlo, Q8, qmeta, Q, d,
                                                          self.cta_group, cnt_gt, self.cand_q, self.cand_e,
                                                          self.count), work)
        else:
            self._timed(lambda: kernels.rank_gemm_count(self.Ehi, self.Elo, self.hi - self.lo, self.lo, Qhi, Qlo, Q,
                                                        d, self.nsplit, tlo, thi, cnt_gt, self.cand_q, self.cand_e,
                                                        self.count), work)


class TransEEval(FilteredRankingEval):
    """skge/run_transe.py:13-29 -- evaluation is L1 whatever ``mdl.l1`` says."""
    model_code = _ext.MODEL_TRANSE


class HolEEval(FilteredRankingEval):
    """skge/run_hole.py:10-19."""
    model_code = _ext.MODEL_HOLE


class RESCALEval(FilteredRankingEval):
    """No evaluator exists in the reference; scores follow skge/rescal.py:31-35."""
    model_code = _ext.MODEL_RESCAL


# ---------------------------------------------------------------------------
# metrics (skge/base.py:1050-1103)
# ---------------------------------------------------------------------------

def compute_scores(pos, hits=10):
    pos = np.asarray(pos)
    mrr = np.mean(1.0 / pos)
    mean_pos = np.mean(pos)
    ans_hits = np.mean(pos <= hits).sum() * 100
    return mrr, mean_pos, ans_hits


def _print_pos(fresult, pos, fpos, epoch, txt):
    mrr, mean_pos, hits = compute_scores(pos)
    fmrr, fmean_pos, fhits = compute_scores(fpos)
    line = "[%3d] %s: MRR = %.2f/%.2f, Mean Rank = %.2f/%.2f, Hits@10 = %.2f/%.2f" % (
        epoch, txt, mrr, fmrr, mean_pos, fmean_pos, hits, fhits)
    log.info(line)
    if fresult is not None:
        fresult.write(line + "\n")
    return fmrr


def ranking_scores(fresult, pos, fpos, epoch, txt):
    hpos = [p for k in pos.keys() for p in pos[k]['head']]
    tpos = [p for k in pos.keys() for p in pos[k]['tail']]
    fhpos = [p for k in fpos.keys() for p in fpos[k]['head']]
    ftpos = [p for k in fpos.keys() for p in fpos[k]['tail']]
    return _print_pos(fresult, np.array(hpos + tpos), np.array(fhpos + ftpos), epoch, txt)


class LinkPredictionEval(object):
    """Area under the precision-recall and ROC curves of ``_scores`` on labelled triples
    (skge/base.py:1034-1047; the metrics come from scikit-learn exactly as in the reference,
    the scores from the device)."""

    def __init__(self, xs, ys):
        ss, os_, ps = list(zip(*xs))
        self.ss, self.ps, self.os, self.ys = list(ss), list(ps), list(os_), ys

    def scores(self, mdl):
        from sklearn.metrics import precision_recall_curve, auc, roc_auc_score
        scores = mdl._scores(self.ss, self.ps, self.os)
        pr, rc, _ = precision_recall_curve(self.ys, scores)
        return auc(rc, pr), roc_auc_score(self.ys, scores)
