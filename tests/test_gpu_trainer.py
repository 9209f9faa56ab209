"""GPU tests of the trainer API (skge/base.py:1195-1427 of the reference): fit(),
minibatching, callbacks, samplers, both execution paths, pickling."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _graph(seed=0, N=300, M=4, T=3000):
    rng = np.random.default_rng(seed)
    tr = np.unique(np.stack([rng.integers(N, size=T), rng.integers(N, size=T), rng.integers(M, size=T)], 1), axis=0)
    return N, M, [tuple(map(int, t)) for t in tr]


@pytest.mark.parametrize('fused', [True, False])
@pytest.mark.parametrize('model', ['transe', 'hole'])
def test_pairwise_fit_reduces_violations(model, fused):
    import skge
    from skge.sample import RandomModeSampler
    N, M, xs = _graph()
    if not fused:
        xs = xs[:600]
    sz = (N, N, M)
    m = skge.TransE(sz, 32) if model == 'transe' else skge.HolE(sz, 32)
    smp = RandomModeSampler(1, [0, 1], xs, sz)
    seen = []

    def cb(trn):
        seen.append((trn.epoch, trn.nviolations, trn.epoch_start > 0, trn.model is m))
        return True

    trn = skge.PairwiseStochasticTrainer(m, nbatches=7, margin=2.0 if model == 'transe' else 0.2,
                                         max_epochs=8 if fused else 3, learning_rate=0.1, samplef=smp.sample,
                                         post_epoch=[cb], fused=fused)
    assert trn._can_fuse('_fused_pair_step') == fused
    trn.fit(xs, np.ones(len(xs)))
    assert [e for e, *_ in seen] == list(range(1, trn.max_epochs + 1))
    assert all(ok and same for _, _, ok, same in seen)
    v = [nv for _, nv, *_ in seen]
    assert 0 < v[-1] < v[0] <= 2 * len(xs)
    assert trn.batch_size == len(xs) // 7
    assert m.margin == trn.model.margin                      # margin lives on the model (base.py:1335)
    nb = np.zeros(N, dtype=int)
    for s, o, _ in xs:
        nb[s] += 1
        nb[o] += 1
    np.testing.assert_array_equal(m.E.neighbours, nb)        # base.py:1364-1367
    if model == 'transe':
        np.testing.assert_allclose(np.linalg.norm(np.asarray(m.E, dtype=np.float64), axis=1)[nb > 0], 1.0,
                                   rtol=1e-5)
        assert sum(m.E.violations) > 0 and sum(m.E.updateCounts) > 0


def test_training_improves_filtered_mrr():
    """End to end on a learnable graph: ranks of held-out triples improve."""
    import skge
    from skge.sample import RandomModeSampler
    from skge.ranking import TransEEval, ranking_scores
    rng = np.random.default_rng(5)
    N, M, d = 200, 3, 16
    Et, Rt = rng.normal(size=(N, d)), rng.normal(size=(M, d))
    triples = []
    for p in range(M):
        for s in range(N):
            o = int(np.argmin(np.abs(Et[s] + Rt[p] - Et).sum(1) + 1e9 * (np.arange(N) == s)))
            triples.append((s, o, p))
    rng.shuffle(triples)
    test, train = triples[:60], triples[60:]
    m = skge.TransE((N, N, M), 32)
    ev = TransEEval(test, triples)
    before = ranking_scores(None, *ev.positions(m), 0, 'TEST')
    smp = RandomModeSampler(1, [0, 1], train, (N, N, M))
    trn = skge.PairwiseStochasticTrainer(m, nbatches=5, margin=2.0, max_epochs=150, learning_rate=0.1,
                                         samplef=smp.sample)
    trn.fit(train, np.ones(len(train)))
    after = ranking_scores(None, *ev.positions(m), 150, 'TEST')
    # the same recipe run by the CPU oracle (float64 numpy, its own random stream)
    from oracle import cpu_oracle as orc
    tr = np.array(train)
    E0 = orc.normalize(rng.uniform(-1, 1, (N, 32)) * np.sqrt(6) / np.sqrt(N + 32))
    R0 = rng.uniform(-1, 1, (M, 32)) * np.sqrt(6) / np.sqrt(M + 32)
    run = orc.PairwiseEpochRunner('transe', E0, R0, tr, (N, N, M), 2.0, 0.1, 5, seed=1)
    for _ in range(150):
        run.epoch()
    _, (cpu_fmrr, _, _) = orc.ranking_scores(*orc.rank_positions('transe', run.E, run.R, np.array(test),
                                                                  np.array(triples)))
    assert after > 1.5 * before and after > 0.05
    assert after > 0.6 * cpu_fmrr, (after, cpu_fmrr)


def test_supplied_negatives_mode_and_callback_break():
    """samplef=None: positives/negatives given in xs/ys (skge/base.py:1350-1357);
    a callback returning False only stops the callback loop (:1289-1291)."""
    import skge
    N, M, xs = _graph(seed=2, T=800)
    rng = np.random.default_rng(0)
    neg = [(int(rng.integers(N)), o, p) for s, o, p in xs] + [(s, int(rng.integers(N)), p) for s, o, p in xs]
    allx = xs + neg
    ys = np.concatenate([np.ones(len(xs)), -np.ones(len(neg))])
    calls = []
    for fused in (True, False):
        m = skge.HolE((N, N, M), 16)
        trn = skge.PairwiseStochasticTrainer(m, nbatches=4, margin=0.2, max_epochs=3, fused=fused,
                                             post_epoch=[lambda t: calls.append(('a', t.epoch)) or False,
                                                         lambda t: calls.append(('b', t.epoch)) or True])
        trn.fit(allx, ys)
        assert trn.epoch == 3 and trn.nviolations > 0
        assert len(trn.pxs) == 2 * len(xs) and len(trn.nxs) == len(neg)
    assert [c for c in calls if c[0] == 'b'] == []            # 'a' returned False -> 'b' never runs
    assert len(calls) == 6


@pytest.mark.parametrize('model', ['hole', 'rescal'])
@pytest.mark.parametrize('fused', [True, False])
def test_logistic_fit_reduces_loss(model, fused):
    import skge
    from skge.param import SGD, AdaGrad
    from skge.sample import RandomModeSampler
    N, M, xs = _graph(seed=1, T=1200 if fused else 400)
    sz = (N, N, M)
    m = skge.HolE(sz, 24) if model == 'hole' else skge.RESCAL(sz, 12)
    smp = RandomModeSampler(1, [0, 1], xs, sz)
    losses = []
    trn = skge.StochasticTrainer(m, nbatches=6, max_epochs=6 if fused else 3, learning_rate=0.1,
                                 samplef=smp.sample, param_update=AdaGrad if model == 'hole' else SGD,
                                 post_epoch=[lambda t: losses.append(t.loss) or True], fused=fused)
    assert trn._can_fuse('_fused_logistic_step') == fused
    trn.fit(xs, np.ones(len(xs)))
    assert len(losses) == trn.max_epochs and losses[-1] < losses[0]
    assert losses[0] == pytest.approx(3 * len(xs) * np.log(2), rel=0.2)   # ~log 2 per example at init


def test_custom_updater_and_sampler_use_the_hook_path():
    """A user-defined param_update / samplef must still work (duck-typed hooks)."""
    import skge
    from skge.param import ParameterUpdate
    N, M, xs = _graph(seed=4, T=300)
    used = {'upd': 0, 'smp': 0}

    class Half(ParameterUpdate):
        def _update(self, g, idx):
            used['upd'] += 1
            self.param[idx] -= 0.5 * self.learning_rate * np.asarray(g)

    def samplef(xys):
        used['smp'] += 1
        return [((x[0], (x[1] + 1) % N, x[2]), -1.0) for x, _ in xys]

    m = skge.TransE((N, N, M), 8)
    trn = skge.PairwiseStochasticTrainer(m, nbatches=3, margin=2.0, max_epochs=2, samplef=samplef,
                                         param_update=Half)
    assert not trn._can_fuse('_fused_pair_step')
    trn.fit(xs, np.ones(len(xs)))
    assert used['upd'] > 0 and used['smp'] == 2 * len(xs)
    nrm = np.linalg.norm(np.asarray(m.E, dtype=np.float64), axis=1)
    np.testing.assert_allclose(nrm, 1.0, rtol=1e-5)           # post-hook still applied after custom updates


def test_model_pickle_round_trip(tmp_path):
    import skge
    from skge.base import Model
    m = skge.HolE((50, 50, 3), 8, rparam=0.1)
    f = str(tmp_path / 'm.pkl')
    m.save(f)
    m2 = Model.load(f)
    assert isinstance(m2, skge.HolE) and m2.rparam == 0.1 and m2.ncomp == 8 and tuple(m2.sz) == (50, 50, 3)
    np.testing.assert_array_equal(np.asarray(m2.E), np.asarray(m.E))
    np.testing.assert_array_equal(np.asarray(m2.R), np.asarray(m.R))
    assert m2.E.post is m.E.post
    s = m2._scores([0, 1], [0, 1], [2, 3])
    np.testing.assert_allclose(s, m._scores([0, 1], [0, 1], [2, 3]))


def test_cli_experiment_flow(tmp_path, monkeypatch):
    """run_transe / run_hole command lines of the reference (README.md:4, run_*_wn18.sh) with the
    non-trident loader: epoch log lines, VALID/TEST ranking lines, best-model pickle at --fout."""
    import pickle
    from skge.run_transe import ExpTransE
    from skge.run_hole import ExpHolE
    from skge.run_rescal import ExpRESCAL
    monkeypatch.chdir(tmp_path)
    rng = np.random.default_rng(0)
    tr = np.unique(np.stack([rng.integers(120, size=2500), rng.integers(120, size=2500), rng.integers(3, size=2500)], 1),
                   axis=0)
    np.savez(tmp_path / 'toy.npz', train=tr)
    fout = str(tmp_path / 'best.pkl')
    ExpTransE().run(['--fin', str(tmp_path / 'toy.npz'), '--test-all', '2', '--nb', '5', '--me', '4',
                     '--margin', '2.0', '--lr', '0.1', '--ncomp', '16', '--fout', fout,
                     '--fgrad', str(tmp_path / 'grad.csv')])
    out = (tmp_path / 'toy-full-kognac-epochs-4-eval-2-margin-2.0.out').read_text()
    assert out.count('violations =') == 5          # 4 epochs + the final with_eval call
    assert 'VALID: MRR =' in out and 'TEST: MRR =' in out and 'Time to fit model' in out
    from skge.base import loads_reference       # reference-format stream (see Model.save)
    st = loads_reference(open(fout, 'rb').read())
    assert set(st) == {'model', 'pos test', 'fpos test', 'pos valid', 'fpos valid', 'exectimes'}
    assert st['model'].E.shape == (120, 16) and len(st['exectimes']) >= 4
    assert (tmp_path / 'grad.csv').read_text().startswith('Entity,Degree,#(violations),#(updates)')
    ExpHolE().run(['--fin', str(tmp_path / 'toy.npz'), '--test-all', '3', '--nb', '5', '--me', '3', '--margin', '0.2',
                   '--lr', '0.1', '--ncomp', '16', '--sampler', 'lcwa'])
    ExpHolE().run(['--fin', str(tmp_path / 'toy.npz'), '--test-all', '3', '--nb', '5', '--me', '2', '--lr', '0.1',
                   '--ncomp', '8', '--no-pairwise'])
    ExpRESCAL().run(['--fin', str(tmp_path / 'toy.npz'), '--test-all', '2', '--nb', '5', '--me', '2', '--lr', '0.1',
                     '--ncomp', '8'])
    with pytest.raises(ValueError, match='Unknown sampler'):
        ExpTransE().run(['--fin', str(tmp_path / 'toy.npz'), '--nb', '5', '--me', '1', '--margin', '1', '--lr', '0.1',
                         '--ncomp', '4', '--sampler', 'bogus'])


def test_link_prediction_eval_auc():
    """skge/base.py:1034-1047."""
    import skge
    from skge.base import LinkPredictionEval
    from sklearn.metrics import roc_auc_score
    N, M, xs = _graph(seed=9, T=400)
    m = skge.HolE((N, N, M), 16)
    rng = np.random.default_rng(0)
    ys = np.where(rng.random(len(xs)) < 0.5, 1, -1)
    pr_auc, roc = LinkPredictionEval(xs, ys).scores(m)
    s, o, p = zip(*xs)
    assert roc == pytest.approx(roc_auc_score(ys, m._scores(s, p, o))) and 0.0 <= pr_auc <= 1.0


def test_fused_step_with_nothing_to_update_and_bad_ids():
    """A minibatch in which no pair violates leaves the parameters untouched (the reference
    returns None and skips _batch_step, skge/base.py:1425-1427); a fully masked minibatch too;
    out-of-range ids raise like the reference's fancy indexing would."""
    import skge
    from skge._modelutil import idx_tensor
    for cls, d in ((skge.TransE, 16), (skge.HolE, 64), (skge.HolE, 20)):
        m = cls((30, 30, 3), d)
        trn = skge.PairwiseStochasticTrainer(m, nbatches=1, margin=-1e6, learning_rate=0.1)
        trn._setup_fused()
        if hasattr(m, '_prepare_fused'):
            m._prepare_fused()
        E0, R0 = np.asarray(m.E).copy(), np.asarray(m.R).copy()
        rng = np.random.default_rng(0)
        pos = tuple(idx_tensor(rng.integers(n, size=50)) for n in (30, 30, 3))
        neg = tuple(idx_tensor(rng.integers(n, size=50)) for n in (30, 30, 3))
        m._fused_pair_step(trn._updaters, pos, neg, None, trn._counts, trn._nviol_dev)
        assert trn._counts.tolist()[:3] == [0, 0, 0] and int(trn._nviol_dev.item()) == 0
        m.margin = 1e6          # everything would violate, but every pair is masked out
        m._fused_pair_step(trn._updaters, pos, neg, torch.zeros(50, dtype=torch.uint8, device='cuda'),
                           trn._counts, trn._nviol_dev)
        assert trn._counts.tolist()[:3] == [0, 0, 0]
        np.testing.assert_array_equal(np.asarray(m.E), E0)
        np.testing.assert_array_equal(np.asarray(m.R), R0)
    from skge.sample import RandomModeSampler
    xs = [(0, 1, 0), (2, 31, 1)]
    smp = RandomModeSampler(1, [0, 1], xs, (30, 30, 3))
    trn = skge.PairwiseStochasticTrainer(skge.TransE((30, 30, 3), 8), nbatches=1, max_epochs=1, samplef=smp.sample)
    with pytest.raises(IndexError):
        trn.fit(xs, np.ones(2))
