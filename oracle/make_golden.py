#!/usr/bin/env python
"""Generate golden vectors from the UNMODIFIED reference (test infrastructure).

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py            # writes tests/golden/*.npz

The reference is imported under the four shims of SURVEY.md section 8c
(trident stub, sys.path for its non-package imports, collections.Hashable,
np.Inf).  It must never share an interpreter with the product package, which
is also importable as ``skge`` -- this script therefore only touches
/root/reference and numpy.  The committed .npz files are what travels to the
GPU box; nothing there reads /root/reference.
"""
import collections
import collections.abc
import importlib.util
import logging
import os
import sys
import types

REF = os.environ.get('SKGE_REFERENCE', '/root/reference')
sys.path[:0] = [REF, os.path.join(REF, 'skge')]
sys.modules['trident'] = types.ModuleType('trident')
collections.Hashable = collections.abc.Hashable
import numpy as np  # noqa: E402
np.Inf = np.inf

import skge  # noqa: E402  (the reference)
import skge.base as base  # noqa: E402
from skge import TransE, HolE, RESCAL, PairwiseStochasticTrainer, StochasticTrainer  # noqa: E402
from skge.param import SGD, AdaGrad  # noqa: E402
from skge import activation_functions as afs  # noqa: E402
from skge.util import ccorr, cconv  # noqa: E402
from skge import sample  # noqa: E402

assert os.path.realpath(skge.__file__).startswith(os.path.realpath(REF)), skge.__file__
logging.disable(logging.CRITICAL)
sys.argv = ['x']


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, 'skge', name + '.py'))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


TransEEval = _load('run_transe').TransEEval
HolEEval = _load('run_hole').HolEEval

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')


def f32(a):
    """fp32-rounded values carried in float64, so input rounding is not
    counted as kernel error."""
    return np.asarray(a, dtype=np.float32).astype(np.float64)


def as_xys(tr, y=1.0):
    return [((int(s), int(o), int(p)), y) for s, o, p in tr]


def set_params(m, **vals):
    for k, v in vals.items():
        getattr(m, k)[...] = v


def rand_triples(rng, n, N, M):
    return np.stack([rng.integers(N, size=n), rng.integers(N, size=n), rng.integers(M, size=n)], 1)


def corrupt(rng, pos, N):
    neg = pos.copy()
    which = rng.integers(2, size=len(pos))
    neg[np.arange(len(pos)), which] = rng.integers(N, size=len(pos))
    return neg


def pairwise_case(name, cls, N, M, d, P, margin, seed, update, scale=1.0, **mkw):
    """One _pairwise_gradients call + one _batch_step of the reference."""
    rng = np.random.default_rng(seed)
    m = cls((N, N, M), d, **mkw)
    E0 = f32(rng.uniform(-1, 1, (N, d)) * scale)
    R0 = f32(rng.uniform(-1, 1, (M, d)) * scale)
    set_params(m, E=E0, R=R0)
    trn = PairwiseStochasticTrainer(m, nbatches=1, margin=margin, max_epochs=1,
                                    learning_rate=0.1, param_update=update,
                                    file_grad=None, file_embed=None)
    pos = rand_triples(rng, P, N, M)
    neg = corrupt(rng, pos, N)
    out = dict(E0=E0, R0=R0, pos=pos, neg=neg, margin=margin, lr=0.1,
               update=update.__name__.lower())
    if cls is TransE:
        out['l1'] = bool(mkw.get('l1', True))
        out['pscores'] = m._scores(pos[:, 0], pos[:, 2], pos[:, 1])
        out['nscores'] = m._scores(neg[:, 0], neg[:, 2], neg[:, 1])
    else:
        out['af'] = m.af.key()
        out['rparam'] = float(m.rparam)
        out['raw_p'] = m._scores(pos[:, 0], pos[:, 2], pos[:, 1])
        out['raw_n'] = m._scores(neg[:, 0], neg[:, 2], neg[:, 1])
    # two consecutive steps so AdaGrad's accumulated p2 is exercised
    for step in (1, 2):
        g = m._pairwise_gradients(as_xys(pos), as_xys(neg))
        out['nviolations%d' % step] = m.nviolations
        if g is not None:
            out['ge%d' % step], out['eidx%d' % step] = np.array(g['E'][0]), np.array(g['E'][1])
            out['gr%d' % step], out['ridx%d' % step] = np.array(g['R'][0]), np.array(g['R'][1])
            trn._batch_step(g)
        out['E%d' % step] = np.array(m.E)
        out['R%d' % step] = np.array(m.R)
    if cls is TransE:
        out['violations'] = np.array(m.E.violations)
    np.savez(os.path.join(OUT, name + '.npz'), **out)
    print(name, 'nviol', out['nviolations1'], out['nviolations2'])


def logistic_case(name, cls, N, M, d, n, seed, update, rparam):
    rng = np.random.default_rng(seed)
    m = cls((N, N, M), d, rparam=rparam)
    E0 = f32(rng.uniform(-1, 1, (N, d)) * 0.5)
    xs = rand_triples(rng, n, N, M)
    ys = np.where(rng.random(n) < 0.4, 1.0, -1.0)
    out = dict(E0=E0, xs=xs, ys=ys, rparam=rparam, lr=0.1, update=update.__name__.lower())
    if cls is RESCAL:
        W0 = f32(rng.uniform(-1, 1, (M, d, d)) * 0.5)
        set_params(m, E=E0, W=W0)
        out['W0'] = W0
        second = 'W'
    else:
        R0 = f32(rng.uniform(-1, 1, (M, d)) * 0.5)
        set_params(m, E=E0, R=R0)
        out['R0'] = R0
        second = 'R'
    trn = StochasticTrainer(m, nbatches=1, max_epochs=1, learning_rate=0.1, param_update=update)
    out['scores'] = np.array(m._scores(xs[:, 0], xs[:, 2], xs[:, 1]))
    xys = [((int(s), int(o), int(p)), float(y)) for (s, o, p), y in zip(xs, ys)]
    for step in (1, 2):
        g = m._gradients(list(xys))
        out['loss%d' % step] = float(m.loss)
        out['ge%d' % step], out['eidx%d' % step] = np.array(g['E'][0]), np.array(g['E'][1])
        out['g2_%d' % step], out['idx2_%d' % step] = np.array(g[second][0]), np.array(g[second][1])
        trn._batch_step(g)
        out['E%d' % step] = np.array(m.E)
        out['P2_%d' % step] = np.array(getattr(m, second))
    np.savez(os.path.join(OUT, name + '.npz'), **out)
    print(name, 'loss', out['loss1'], out['loss2'])


def ranking_case(name, cls, Ev, N, M, d, ntrue, ntest, seed):
    rng = np.random.default_rng(seed)
    m = cls((N, N, M), d)
    E0 = f32(rng.uniform(-1, 1, (N, d)) * 0.5)
    R0 = f32(rng.uniform(-1, 1, (M, d)) * 0.5)
    set_params(m, E=E0, R=R0)
    true = np.unique(rand_triples(rng, ntrue, N, M), axis=0)
    test = true[rng.choice(len(true), ntest, replace=False)]
    ev = Ev([tuple(map(int, t)) for t in test], [tuple(map(int, t)) for t in true], -1)
    pos, fpos = ev.positions(m)
    rel = np.array(list(pos.keys()))
    flat = lambda dct, k: np.concatenate([np.array(dct[p][k]) for p in rel])  # noqa: E731
    cnt = np.array([len(pos[p]['head']) for p in rel])
    mrr = base.compute_scores(np.concatenate([flat(pos, 'head'), flat(pos, 'tail')]))
    fmrr = base.compute_scores(np.concatenate([flat(fpos, 'head'), flat(fpos, 'tail')]))
    np.savez(os.path.join(OUT, name + '.npz'), E0=E0, R0=R0, true=true, test=test,
             rel=rel, cnt=cnt, pos_head=flat(pos, 'head'), pos_tail=flat(pos, 'tail'),
             fpos_head=flat(fpos, 'head'), fpos_tail=flat(fpos, 'tail'),
             raw=np.array(mrr), filt=np.array(fmrr))
    print(name, 'mrr', mrr[0], fmrr[0])


def appendix_a():
    """The known-answer case of SURVEY.md Appendix A, regenerated."""
    E0 = (np.arange(20).reshape(5, 4) % 7 - 3) / 4.0
    R0 = (np.arange(8).reshape(2, 4) % 5 - 2) / 4.0
    W0 = (np.arange(32).reshape(2, 4, 4) % 9 - 4) / 8.0
    pos = np.array([(0, 1, 0), (0, 1, 0), (2, 3, 1), (2, 3, 1)])
    neg = np.array([(4, 1, 0), (0, 2, 0), (1, 3, 1), (2, 0, 1)])
    out = dict(E0=E0, R0=R0, W0=W0, pos=pos, neg=neg)
    # A1 TransE
    m = TransE((5, 5, 2), 4, l1=True)
    set_params(m, E=E0, R=R0)
    trn = PairwiseStochasticTrainer(m, nbatches=1, margin=2.0, learning_rate=0.1,
                                    file_grad=None, file_embed=None)
    g = m._pairwise_gradients(as_xys(pos), as_xys(neg))
    trn._batch_step(g)
    out.update(a1_ge=g['E'][0], a1_eidx=g['E'][1], a1_gr=g['R'][0], a1_ridx=g['R'][1],
               a1_E=np.array(m.E), a1_R=np.array(m.R), a1_nviol=m.nviolations)
    # A2 HolE
    m = HolE((5, 5, 2), 4, rparam=0.0, af=afs['sigmoid'])
    set_params(m, E=E0, R=R0)
    trn = PairwiseStochasticTrainer(m, nbatches=1, margin=0.2, learning_rate=0.1,
                                    file_grad=None, file_embed=None)
    out['a2_raw_p'] = m._scores(pos[:, 0], pos[:, 2], pos[:, 1])
    out['a2_raw_n'] = m._scores(neg[:, 0], neg[:, 2], neg[:, 1])
    g = m._pairwise_gradients(as_xys(pos), as_xys(neg))
    trn._batch_step(g)
    out.update(a2_ge=g['E'][0], a2_eidx=g['E'][1], a2_gr=g['R'][0], a2_ridx=g['R'][1],
               a2_E=np.array(m.E), a2_R=np.array(m.R), a2_nviol=m.nviolations)
    # A3 RESCAL logistic
    m = RESCAL((5, 5, 2), 4, rparam=0.0)
    set_params(m, E=E0, W=W0)
    xs = np.array([pos[0], pos[2], neg[0], neg[1], neg[2], neg[3]])
    ys = np.array([1, 1, -1, -1, -1, -1], dtype=np.float64)
    g = m._gradients([((int(s), int(o), int(p)), float(y)) for (s, o, p), y in zip(xs, ys)])
    out.update(a3_xs=xs, a3_ys=ys, a3_loss=float(m.loss), a3_ge=g['E'][0], a3_eidx=g['E'][1],
               a3_gw=g['W'][0], a3_pidx=g['W'][1])
    # A4 ranking
    true = np.array([(0, 1, 0), (0, 2, 0), (2, 3, 1), (4, 3, 1), (2, 0, 1), (1, 4, 0)])
    test = np.array([(0, 1, 0), (2, 3, 1), (1, 4, 0)])
    out.update(a4_true=true, a4_test=test)
    for tag, cls, Ev in (('te', TransE, TransEEval), ('ho', HolE, HolEEval)):
        m = cls((5, 5, 2), 4)
        set_params(m, E=E0, R=R0)
        ev = Ev([tuple(map(int, t)) for t in test], [tuple(map(int, t)) for t in true], -1)
        p, fp = ev.positions(m)
        for k in p:
            for side in ('head', 'tail'):
                out['a4_%s_pos_%d_%s' % (tag, k, side)] = np.array(p[k][side])
                out['a4_%s_fpos_%d_%s' % (tag, k, side)] = np.array(fp[k][side])
    np.savez(os.path.join(OUT, 'appendix_a.npz'), **out)
    print('appendix_a ok')


def primitives():
    rng = np.random.default_rng(5)
    out = {}
    for d in (4, 7, 50, 150, 256):
        a, b = rng.normal(size=(3, d)), rng.normal(size=(3, d))
        out['a%d' % d], out['b%d' % d] = a, b
        out['ccorr%d' % d], out['cconv%d' % d] = ccorr(a, b), cconv(a, b)
    x = rng.normal(size=(6, 5))
    out['x'] = x
    for k, af in afs.items():
        if k == 'softplus':
            continue
        fx = af.f(x[:, 0])
        out['af_f_' + k], out['af_g_' + k] = fx, af.g_given_f(fx)
    from skge.param import normalize, normless1
    out['normalize_all'] = normalize(x.copy())
    out['normalize_idx'] = normalize(x.copy() * 3, np.array([1, 4]))
    out['normless1_all'] = normless1(x.copy())
    out['normless1_idx'] = normless1(x.copy() * 3, np.array([0, 2, 5]))
    out['normless1_idx_small'] = normless1(x.copy() * 0.1, np.array([0, 2, 5]))
    np.savez(os.path.join(OUT, 'primitives.npz'), **out)
    print('primitives ok')


def sampler_case():
    """RandomModeSampler invariants (not stream parity): every emitted negative
    differs from its positive in exactly the sampled slot and is not a training
    triple; saturated slots are skipped after ntries."""
    rng = np.random.default_rng(3)
    N, M = 6, 2
    xs = [tuple(map(int, t)) for t in np.unique(rand_triples(rng, 40, N, M), axis=0)]
    smp = sample.RandomModeSampler(1, [0, 1], xs, (N, N, M))
    np.random.seed(0)
    res = smp.sample([(x, 1.0) for x in xs])
    np.savez(os.path.join(OUT, 'sampler.npz'), xs=np.array(xs), N=N, M=M,
             neg=np.array([t for t, _ in res]), nneg=len(res))
    print('sampler', len(xs), len(res))


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    appendix_a()
    primitives()
    sampler_case()
    pairwise_case('transe_l1_adagrad', TransE, 60, 5, 50, 48, 2.0, 11, AdaGrad, l1=True)
    pairwise_case('transe_l2_sgd', TransE, 60, 5, 16, 48, 1.0, 12, SGD, scale=0.5, l1=False)
    pairwise_case('transe_l1_sgd_d7', TransE, 30, 3, 7, 64, 2.0, 13, SGD, l1=True)
    pairwise_case('hole_sigmoid_adagrad', HolE, 60, 5, 150, 48, 0.2, 21, AdaGrad, scale=0.3,
                  af=afs['sigmoid'])
    pairwise_case('hole_tanh_sgd_rparam', HolE, 40, 4, 16, 64, 0.5, 22, SGD, scale=0.5,
                  af=afs['tanh'], rparam=0.1)
    pairwise_case('hole_linear_adagrad_d256', HolE, 50, 4, 256, 32, 1.0, 23, AdaGrad, scale=0.2,
                  af=afs['linear'])
    pairwise_case('hole_relu_sgd_d10', HolE, 30, 3, 10, 64, 0.3, 24, SGD, scale=0.7, af=afs['relu'])
    # config 2's row length: 150 = 2 * 3 * 5 * 5 (mixed-radix transform on the device)
    pairwise_case('hole_sigmoid_adagrad_d150', HolE, 60, 5, 150, 48, 0.2, 25, AdaGrad, scale=0.3,
                  af=afs['sigmoid'])
    logistic_case('hole_logistic_adagrad', HolE, 50, 4, 32, 90, 31, AdaGrad, 0.05)
    logistic_case('hole_logistic_sgd_d150', HolE, 50, 4, 150, 60, 32, SGD, 0.0)
    logistic_case('rescal_logistic_sgd', RESCAL, 40, 4, 20, 90, 33, SGD, 0.0)
    logistic_case('rescal_logistic_adagrad_rparam', RESCAL, 40, 3, 12, 70, 34, AdaGrad, 0.1)
    ranking_case('rank_transe', TransE, TransEEval, 80, 4, 20, 400, 40, 41)
    ranking_case('rank_hole', HolE, HolEEval, 80, 4, 24, 400, 40, 42)
