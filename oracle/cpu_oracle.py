"""CPU oracle for the scikit-kge hot path -- TEST INFRASTRUCTURE ONLY.

This file is a float64 numpy restatement of the reference's algorithms for the
embedding-training / filtered-ranking hot path.  It is *not* part of the
product: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline``
/ ``--impl reference`` legs of ``bench.py`` may import it.  The product path
(``scikit-kge_b200/skge``) never imports anything from ``oracle/``.

Parity pinning: the reference ships no tests and no golden vectors (SURVEY.md
section 4), so this restatement is pinned against outputs of the reference
itself, generated in the build container by ``oracle/make_golden.py`` (which
imports /root/reference under four import shims) and committed under
``tests/golden/``.  ``tests/test_oracle_golden.py`` checks every function
below against those files.

Conventions (all cited lines are relative to /root/reference):
  * triples are int arrays of shape (n, 3) in the reference's (s, o, p) column
    order (skge/util.py:104-110, skge/base.py:509-511);
  * parameters are float64 ndarrays: E (N, d), R (M, d), W (M, d, d);
  * per-row gradients are MEANS over the occurrences of a row in the minibatch
    (skge/util.py:97-101 and every ``Sm.dot(...) / n`` call site).
"""
import numpy as np

# --------------------------------------------------------------------------
# numeric primitives (skge/util.py)
# --------------------------------------------------------------------------


def cconv(a, b):
    """Circular convolution c_k = sum_i a_i b_{(k-i) mod d}  (skge/util.py:8-27).

    Computed, as the reference does, through the complex FFT of the last axis.
    """
    return np.fft.ifft(np.fft.fft(a) * np.fft.fft(b)).real


def ccorr(a, b):
    """Circular correlation c_k = sum_i a_i b_{(i+k) mod d}  (skge/util.py:30-50)."""
    return np.fft.ifft(np.conj(np.fft.fft(a)) * np.fft.fft(b)).real


def ccorr_direct(a, b):
    """O(d^2) definition of ccorr; used to cross-check the FFT form."""
    a = np.atleast_2d(a)
    b = np.atleast_2d(b)
    d = a.shape[-1]
    k = np.arange(d)
    out = np.zeros(np.broadcast_shapes(a.shape, b.shape))
    for i in range(d):
        out += a[..., i:i + 1] * b[..., (i + k) % d]
    return out


def cconv_direct(a, b):
    """O(d^2) definition of cconv; used to cross-check the FFT form."""
    a = np.atleast_2d(a)
    b = np.atleast_2d(b)
    d = a.shape[-1]
    k = np.arange(d)
    out = np.zeros(np.broadcast_shapes(a.shape, b.shape))
    for i in range(d):
        out += a[..., i:i + 1] * b[..., (k - i) % d]
    return out


def segment_mean(idx, rows):
    """Mean of ``rows`` grouped by ``idx``; returns (sorted unique ids, means).

    Restates ``grad_sum_matrix`` + ``Sm.dot(G) / n`` (skge/util.py:53-101 and
    e.g. skge/transe.py:128-136): np.unique gives the ascending unique ids and
    the inverse map, the CSR selector sums the rows of each id, ``n`` is the
    number of occurrences.
    """
    idx = np.asarray(idx, dtype=np.int64)
    uidx, inv = np.unique(idx, return_inverse=True)
    acc = np.zeros((len(uidx),) + rows.shape[1:], dtype=np.float64)
    np.add.at(acc, inv, rows)
    n = np.bincount(inv, minlength=len(uidx)).astype(np.float64)
    return uidx, acc / n.reshape((-1,) + (1,) * (rows.ndim - 1))


# --------------------------------------------------------------------------
# activation functions (skge/actfun.py:13-57)
# --------------------------------------------------------------------------


def af_f(name, x):
    if name == 'linear':
        return x
    if name == 'sigmoid':
        return 1.0 / (1.0 + np.exp(-x))
    if name == 'tanh':
        return np.tanh(x)
    if name == 'relu':
        return np.maximum(0, x)
    raise ValueError(name)


def af_g_given_f(name, fx):
    if name == 'linear':
        return np.ones(fx.shape[0])
    if name == 'sigmoid':
        return fx * (1.0 - fx)
    if name == 'tanh':
        return 1 - fx ** 2
    if name == 'relu':
        return (fx > 0).astype(np.float64)
    raise ValueError(name)


# --------------------------------------------------------------------------
# parameter post-processing and updaters (skge/param.py:108-174)
# --------------------------------------------------------------------------


def normalize(M, idx=None):
    """Unit-L2 rows, all rows or only ``idx`` (skge/param.py:161-167)."""
    if idx is None:
        return M / np.sqrt(np.sum(M ** 2, axis=1))[:, None]
    nrm = np.sqrt(np.sum(M[idx] ** 2, axis=1))[:, None]
    M[idx] = M[idx] / nrm
    return M


def normless1(M, idx=None):
    """Divide rows by max(1, squared norm) (skge/param.py:170-174).

    With ``idx=None`` the reference's ``M[None]`` adds a leading axis, so the
    sum runs over ROWS and each COLUMN is divided by max(1, sum_rows x^2)
    (quirk kept; it only matters at parameter creation).
    """
    if idx is None:
        nrm = np.sum(M ** 2, axis=0)[None, :]
        nrm[nrm < 1] = 1
        return M / nrm
    nrm = np.sum(M[idx] ** 2, axis=1)[:, None]
    nrm[nrm < 1] = 1
    M[idx] = M[idx] / nrm
    return M


POSTS = {None: None, 'normalize': normalize, 'normless1': normless1}


def sgd_update(param, g, idx, lr, post=None):
    """param[idx] -= lr*g, then post on idx (skge/param.py:115-118,129-130)."""
    param[idx] -= lr * g
    if post is not None:
        POSTS[post](param, idx)
    return param


def adagrad_update(param, p2, g, idx, lr, post=None):
    """AdaGrad row update (skge/param.py:140-155) followed by the post hook."""
    p2[idx] += g * g
    H = np.maximum(np.sqrt(p2[idx]), 1e-7)
    param[idx] -= lr * g / H
    if post is not None:
        POSTS[post](param, idx)
    return param, p2


# --------------------------------------------------------------------------
# TransE (skge/transe.py)
# --------------------------------------------------------------------------


def transe_scores(E, R, ss, ps, os, l1=True):
    """-||E[s]+R[p]-E[o]||_1, or minus the SQUARED L2 distance (skge/transe.py:25-46)."""
    diff = E[ss] + R[ps] - E[os]
    if l1:
        return -np.sum(np.abs(diff), axis=1)
    return -np.sum(diff ** 2, axis=1)


def transe_pairwise_gradients(E, R, pos, neg, margin, l1=True):
    """Pairwise-margin gradients of TransE (skge/transe.py:48-165).

    ``pos`` / ``neg`` are (P, 3) arrays in (s, o, p) order, paired row by row.
    Returns (grads, info): grads is None when no pair violates
    (skge/transe.py:90-91), else {'E': (ge, eidx), 'R': (gr, ridx)}; info holds
    pscores / nscores / the violating pair indices and the per-entity violation
    increments of skge/transe.py:78-83.
    """
    sp, op, pp = pos[:, 0], pos[:, 1], pos[:, 2]
    sn, on, pn = neg[:, 0], neg[:, 1], neg[:, 2]
    pscores = transe_scores(E, R, sp, pp, op, l1)
    nscores = transe_scores(E, R, sn, pn, on, l1)
    ind = np.where(nscores + margin > pscores)[0]          # transe.py:73
    viol_inc = np.zeros(E.shape[0], dtype=np.int64)
    for i in ind:                                           # transe.py:78-83
        for u in {sn[i], on[i], sp[i], op[i]}:
            viol_inc[u] += 1
    info = dict(pscores=pscores, nscores=nscores, ind=ind, nviolations=len(ind),
                violations=viol_inc)
    if len(ind) == 0:
        return None, info
    sp, op, pp, sn, on, pn = sp[ind], op[ind], pp[ind], sn[ind], on[ind], pn[ind]
    pg = E[op] - R[pp] - E[sp]                              # transe.py:103-104
    ng = E[on] - R[pn] - E[sn]
    if l1:
        pg = np.sign(-pg)                                   # transe.py:115-117
        ng = np.sign(ng)
    else:
        pg = -pg                                            # transe.py:120-121
    eidx, ge = segment_mean(np.concatenate((sp, op, sn, on)),
                            np.vstack((pg, -pg, ng, -ng)))  # transe.py:128-136
    ridx, gr = segment_mean(np.concatenate((pp, pn)), np.vstack((pg, ng)))  # :158-160
    return {'E': (ge, eidx), 'R': (gr, ridx)}, info


# --------------------------------------------------------------------------
# HolE (skge/hole.py)
# --------------------------------------------------------------------------


def hole_scores(E, R, ss, ps, os):
    """sum_k R[p]_k ccorr(E[s], E[o])_k  (skge/hole.py:19-20)."""
    return np.sum(R[ps] * ccorr(E[ss], E[os]), axis=1)


def hole_gradients(E, R, xs, ys, rparam=0.0):
    """Logistic-loss gradients of HolE (skge/hole.py:22-42).

    Returns (grads, loss)."""
    ss, os, ps = xs[:, 0], xs[:, 1], xs[:, 2]
    ys = np.asarray(ys, dtype=np.float64)
    yscores = ys * hole_scores(E, R, ss, ps, os)
    loss = np.sum(np.logaddexp(0, -yscores))
    fs = -(ys * af_f('sigmoid', -yscores))[:, None]
    ridx, gr = segment_mean(ps, fs * ccorr(E[ss], E[os]))
    gr = gr + rparam * R[ridx]
    eidx, ge = segment_mean(np.concatenate((ss, os)),
                            np.vstack((fs * ccorr(R[ps], E[os]),
                                       fs * cconv(E[ss], R[ps]))))
    ge = ge + rparam * E[eidx]
    return {'E': (ge, eidx), 'R': (gr, ridx)}, loss


def hole_pairwise_gradients(E, R, pos, neg, margin, af='sigmoid', rparam=0.0):
    """Pairwise-margin gradients of HolE (skge/hole.py:44-100).

    rparam is applied to the relation rows only (hole.py:83; :98 is commented
    out in the reference).  The two discarded timing FFT batches
    (hole.py:71-73, 88-90) are not reproduced.
    """
    sp, op, pp = pos[:, 0], pos[:, 1], pos[:, 2]
    sn, on, pn = neg[:, 0], neg[:, 1], neg[:, 2]
    raw_p = hole_scores(E, R, sp, pp, op)
    raw_n = hole_scores(E, R, sn, pn, on)
    pscores = af_f(af, raw_p)
    nscores = af_f(af, raw_n)
    ind = np.where(nscores + margin > pscores)[0]           # hole.py:56
    info = dict(raw_p=raw_p, raw_n=raw_n, pscores=pscores, nscores=nscores,
                ind=ind, nviolations=len(ind))
    if len(ind) == 0:
        return None, info
    sp, op, pp, sn, on, pn = sp[ind], op[ind], pp[ind], sn[ind], on[ind], pn[ind]
    gp = -af_g_given_f(af, pscores[ind])[:, None]           # hole.py:66-67
    gn = af_g_given_f(af, nscores[ind])[:, None]
    ridx, gr = segment_mean(np.concatenate((pp, pn)),
                            np.vstack((gp * ccorr(E[sp], E[op]),
                                       gn * ccorr(E[sn], E[on]))))
    gr = gr + rparam * R[ridx]                              # hole.py:82-83
    eidx, ge = segment_mean(np.concatenate((sp, sn, op, on)),
                            np.vstack((gp * ccorr(R[pp], E[op]),
                                       gn * ccorr(R[pn], E[on]),
                                       gp * cconv(E[sp], R[pp]),
                                       gn * cconv(E[sn], R[pn]))))  # hole.py:86-97
    return {'E': (ge, eidx), 'R': (gr, ridx)}, info


# --------------------------------------------------------------------------
# RESCAL (skge/rescal.py)
# --------------------------------------------------------------------------


def rescal_scores(E, W, ss, ps, os):
    """E[s]^T W[p] E[o]  (skge/rescal.py:31-35)."""
    return np.einsum('ni,nij,nj->n', E[ss], W[ps], E[os])


def rescal_gradients(E, W, xs, ys, rparam=0.0):
    """Logistic-loss gradients of RESCAL (skge/rescal.py:37-76).

    Returns (grads, loss) with grads = {'E': (ge, eidx), 'W': (gw, pidx)}.
    The memoisation of rescal.py:43-50 is a cache only and is not restated.
    """
    ss, os, ps = xs[:, 0], xs[:, 1], xs[:, 2]
    ys = np.asarray(ys, dtype=np.float64)
    EW = np.einsum('ni,nij->nj', E[ss], W[ps])              # E[s]^T W[p]
    WE = np.einsum('nij,nj->ni', W[ps], E[os])              # W[p] E[o]
    yscores = ys * np.sum(E[ss] * WE, axis=1)
    loss = np.sum(np.logaddexp(0, -yscores))
    fs = -(ys * af_f('sigmoid', -yscores))[:, None]
    pidx = np.unique(ps)
    d = E.shape[1]
    gw = np.zeros((len(pidx), d, d))
    for i, p in enumerate(pidx):                            # rescal.py:61-70
        ind = np.where(ps == p)[0]
        gw[i] = E[ss[ind]].T.dot(fs[ind] * E[os[ind]]) / len(ind)
        gw[i] += rparam * W[p]
    eidx, ge = segment_mean(np.concatenate((ss, os)), np.vstack((fs * WE, fs * EW)))
    ge = ge + rparam * E[eidx]
    return {'E': (ge, eidx), 'W': (gw, pidx)}, loss


# --------------------------------------------------------------------------
# negative sampling (skge/sample.py:10-46)
# --------------------------------------------------------------------------


def random_mode_sample(rng, xs_set, x, sz, modes=(0, 1), n=1, ntries=100):
    """Negatives of one positive ``x`` = (s, o, p): for each of ``n`` rounds and
    each mode, up to ``ntries`` uniform draws of slot ``mode`` rejected against
    the training-triple set (skge/sample.py:17-25, 38-46).  Returns a list of
    (s, o, p) tuples (those that exhausted their tries are skipped)."""
    out = []
    for _ in range(n):
        for mode in modes:
            nex = list(x)
            for _ in range(ntries):
                nex[mode] = int(rng.integers(sz[mode]))
                if tuple(nex) not in xs_set:
                    out.append(tuple(nex))
                    break
    return out


# --------------------------------------------------------------------------
# filtered ranking (skge/base.py:739-759, 913-1031; run_transe.py, run_hole.py)
# --------------------------------------------------------------------------


def _eval_hooks(kind, E, RW):
    """(prepare, scores_o, scores_s) of the per-model evaluators.

    transe: skge/run_transe.py:13-29 (always L1); hole: skge/run_hole.py:10-19;
    rescal: no reference evaluator exists -- derived from skge/rescal.py:31-35.
    """
    st = {}
    if kind == 'transe':
        def prepare(p):
            st['ER'] = E + RW[p]
        def scores_o(s, p):
            return -np.sum(np.abs(st['ER'][s] - E), axis=1)
        def scores_s(o, p):
            return -np.sum(np.abs(st['ER'] - E[o]), axis=1)
    elif kind == 'hole':
        def prepare(p):
            st['ER'] = ccorr(RW[p], E)
        def scores_o(s, p):
            return np.dot(st['ER'], E[s])
        def scores_s(o, p):
            return np.dot(E, st['ER'][o])
    elif kind == 'rescal':
        def prepare(p):
            st['W'] = RW[p]
        def scores_o(s, p):
            return E.dot(st['W'].T.dot(E[s]))
        def scores_s(o, p):
            return E.dot(st['W'].dot(E[o]))
    else:
        raise ValueError(kind)
    return prepare, scores_o, scores_s


def build_filter_index(test, true_triples):
    """idx[p] = [(s, o)...] over test; tt[p]['os'][s] / tt[p]['ss'][o] over all
    true triples (skge/base.py:739-752)."""
    idx, tt = {}, {}
    for s, o, p in np.asarray(test).tolist():
        idx.setdefault(p, []).append((s, o))
    for s, o, p in np.asarray(true_triples).tolist():
        e = tt.setdefault(p, {'ss': {}, 'os': {}})
        e['os'].setdefault(s, []).append(o)
        e['ss'].setdefault(o, []).append(s)
    return idx, tt


def rank_positions(kind, E, RW, test, true_triples, tie='argsort', with_scores=False):
    """Raw and filtered ranks of every test triple in both directions
    (skge/base.py:913-1031).

    tie='argsort' reproduces the reference literally (1-based position in
    ``argsort(scores)[::-1]``); tie='count' uses rank = 1 + #{score > target},
    which is identical whenever the target's score is not tied.
    Returns (pos, fpos[, margins]) in the reference's layout
    ``{p: {'head': [...], 'tail': [...]}}``; ``margins`` (optional) records per
    query the smallest |score - target score| over the other entities, so
    tests can excuse reference ties within 1e-6.
    """
    idx, tt = build_filter_index(test, true_triples)
    prepare, scores_o, scores_s = _eval_hooks(kind, E, RW)

    def rank_of(scores, target):
        if tie == 'argsort':
            order = np.argsort(scores)[::-1]
            return int(np.where(order == target)[0][0]) + 1
        return 1 + int(np.sum(scores > scores[target]))

    def margin_of(scores, target):
        dlt = np.abs(scores - scores[target])
        dlt[target] = np.inf
        return float(dlt.min())

    pos, fpos, margins = {}, {}, {}
    for p, sos in idx.items():
        ppos = {'head': [], 'tail': []}
        pfpos = {'head': [], 'tail': []}
        pm = {'head': [], 'tail': []}
        prepare(p)
        for s, o in sos:
            sc = scores_o(s, p).flatten()
            ppos['tail'].append(rank_of(sc, o))
            pm['tail'].append(margin_of(sc, o))
            rm = [i for i in tt[p]['os'][s] if i != o]
            sc[rm] = -np.inf
            pfpos['tail'].append(rank_of(sc, o))
            sc = scores_s(o, p).flatten()
            ppos['head'].append(rank_of(sc, s))
            pm['head'].append(margin_of(sc, s))
            rm = [i for i in tt[p]['ss'][o] if i != s]
            sc[rm] = -np.inf
            pfpos['head'].append(rank_of(sc, s))
        pos[p], fpos[p], margins[p] = ppos, pfpos, pm
    if with_scores:
        return pos, fpos, margins
    return pos, fpos


def compute_scores(pos, hits=10):
    """MRR, mean rank, Hits@k in percent (skge/base.py:1099-1103)."""
    pos = np.asarray(pos)
    return np.mean(1.0 / pos), np.mean(pos), np.mean(pos <= hits).sum() * 100


def ranking_scores(pos, fpos, hits=10):
    """Head+tail concatenation over relations (skge/base.py:1050-1059)."""
    def flat(d):
        return np.array([r for k in d for r in d[k]['head']] +
                        [r for k in d for r in d[k]['tail']])
    return compute_scores(flat(pos), hits), compute_scores(flat(fpos), hits)


# --------------------------------------------------------------------------
# trainer loops (skge/base.py:1236-1316, 1348-1427) -- used by the CPU baseline
# --------------------------------------------------------------------------


def batch_slices(n, nbatches):
    """Minibatch boundaries of ``_optim``: nbatches slices of n//nbatches plus a
    remainder slice when n % nbatches != 0 (skge/base.py:1246-1252, 1268)."""
    bs = n // nbatches
    cuts = np.arange(bs, n, bs)
    return np.split(np.arange(n), cuts)


class PairwiseEpochRunner(object):
    """One epoch of PairwiseStochasticTrainer with a RandomModeSampler and
    AdaGrad/SGD (skge/base.py:1254-1291, 1394-1427).  Plain numpy/python like
    the reference; serves as the timed CPU baseline ("port")."""

    def __init__(self, kind, E, R, xs, sz, margin, lr, nbatches, opt='adagrad',
                 l1=True, af='sigmoid', rparam=0.0, seed=0):
        self.kind, self.E, self.R = kind, E, R
        self.xs = np.asarray(xs)
        self.xs_set = set(map(tuple, self.xs.tolist()))
        self.sz, self.margin, self.lr, self.nb = sz, margin, lr, nbatches
        self.opt, self.l1, self.af, self.rparam = opt, l1, af, rparam
        self.rng = np.random.default_rng(seed)
        self.p2E, self.p2R = np.zeros_like(E), np.zeros_like(R)
        self.post = 'normalize' if kind == 'transe' else 'normless1'
        self.nviolations = 0

    def process_batch(self, batch):
        pos, neg = [], []
        for x in self.xs[batch].tolist():
            for nx in random_mode_sample(self.rng, self.xs_set, tuple(x), self.sz):
                pos.append(x)
                neg.append(nx)
        pos, neg = np.array(pos), np.array(neg)
        if self.kind == 'transe':
            g, info = transe_pairwise_gradients(self.E, self.R, pos, neg, self.margin, self.l1)
        else:
            g, info = hole_pairwise_gradients(self.E, self.R, pos, neg, self.margin,
                                              self.af, self.rparam)
        if g is None:
            return
        self.nviolations += info['nviolations']
        for (param, p2, key, post) in ((self.E, self.p2E, 'E', self.post),
                                       (self.R, self.p2R, 'R', None)):
            gg, idx = g[key]
            if self.opt == 'adagrad':
                adagrad_update(param, p2, gg, idx, self.lr, post)
            else:
                sgd_update(param, gg, idx, self.lr, post)

    def epoch(self, max_batches=None):
        idx = self.rng.permutation(len(self.xs))
        self.nviolations = 0
        done = 0
        for bi, batch in enumerate(np.split(idx, np.arange(len(idx) // self.nb, len(idx),
                                                            len(idx) // self.nb))):
            if max_batches is not None and bi >= max_batches:
                break
            self.process_batch(batch)
            done += len(batch)
        return done
