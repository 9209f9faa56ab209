"""Config-sized end-to-end parity ("end-to-end metric parity (MRR / Hits) on a
WN18-shaped synthetic graph").  BASELINE configs 1 and 2 exactly -- WN18 shape (40,943 entities,
18 relations, 141,442 training triples), nb = 100, AdaGrad lr 0.1; TransE d = 50 L1 margin 2.0 for
5 epochs, HolE d = 150 sigmoid margin 0.2 for 1 epoch -- run through PairwiseStochasticTrainer.fit
on the GPU and through the oracle's float64 restatement of the same loop
(skge/base.py:1254-1291, 1348-1427), with SUPPLIED negatives and scripted shuffles so that both
sides see identical minibatches.  Then the filtered ranking of test triples: on the oracle's final
parameters the GPU ranks must equal the oracle's, and the metrics of the two trained models must
agree.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _load_synth():
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('skge_b200_synth_t', os.path.join(root, 'scikit-kge_b200', 'skge',
                                                                                    'synth.py'))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _workload(n_test):
    synth = _load_synth()
    g = synth.make_graph('wn18', device='cpu')
    N, M = g['N'], g['M']
    train = g['train'].numpy()
    rng = np.random.default_rng(11)
    neg = train.copy()                      # one corruption per positive: subject or object replaced
    side = rng.integers(2, size=len(train))
    neg[np.arange(len(train)), side] = rng.integers(N, size=len(train))
    true = np.concatenate([train, g['valid'].numpy(), g['test'].numpy()])
    return N, M, train, neg, true, g['test'].numpy()[:n_test]


def _oracle_fit(orc, kind, E, R, P, Nn, perms, nb, margin, lr, af='sigmoid'):
    """The supplied-negatives loop of skge/base.py:1350-1357, 1254-1291, 1390-1427 in float64."""
    n = len(P)
    p2E, p2R = np.zeros_like(E), np.zeros_like(R)
    bs = n // nb
    cuts = list(range(bs, n, bs))
    viol = []
    for perm_p, perm_n, perm_b in perms:
        P, Nn = P[perm_p], Nn[perm_n]       # independent shuffles of positives and negatives (:1390-1392)
        nv = 0
        for batch in np.split(perm_b, cuts):
            pos, neg = P[batch], Nn[batch]
            if kind == 'transe':
                g, info = orc.transe_pairwise_gradients(E, R, pos, neg, margin, True)
            else:
                g, info = orc.hole_pairwise_gradients(E, R, pos, neg, margin, af, 0.0)
            nv += info['nviolations']
            if g is None:
                continue
            orc.adagrad_update(E, p2E, g['E'][0], g['E'][1], lr, 'normalize' if kind == 'transe' else 'normless1')
            orc.adagrad_update(R, p2R, g['R'][0], g['R'][1], lr, None)
        viol.append(nv)
    return viol


@pytest.mark.parametrize('kind,d,margin,epochs,n_test', [('transe', 50, 2.0, 5, 600), ('hole', 150, 0.2, 1, 300)])
def test_config_sized_training_and_ranking_match_the_oracle(kind, d, margin, epochs, n_test):
    import skge
    from skge.param import AdaGrad
    from skge.ranking import TransEEval, HolEEval
    from oracle import cpu_oracle as orc
    synth = _load_synth()
    N, M, train, neg, true, test = _workload(n_test)
    nb, lr = 100, 0.1
    E0, R0 = synth.init_embeddings(kind, N, M, d, device='cpu')
    n = len(train)
    rng = np.random.default_rng(3)
    perms = [tuple(rng.permutation(n) for _ in range(3)) for _ in range(epochs)]

    # --- GPU: the public trainer API, supplied-negatives mode, scripted shuffles -----------------
    cls = skge.TransE if kind == 'transe' else skge.HolE
    m = cls((N, N, M), d)
    m.E.data.copy_(E0)
    m.R.data.copy_(R0)
    xs = [tuple(t) for t in train.tolist()] + [tuple(t) for t in neg.tolist()]
    ys = np.concatenate([np.ones(n), -np.ones(n)])
    viol_gpu = []
    trn = skge.PairwiseStochasticTrainer(m, nbatches=nb, margin=margin, max_epochs=epochs, learning_rate=lr,
                                         param_update=AdaGrad,
                                         post_epoch=[lambda t: viol_gpu.append(t.nviolations) or True])
    script = iter([p for ep in perms for p in ep])
    trn._randperm = lambda k: torch.from_numpy(next(script)).to(m.E.data.device)
    trn.fit(xs, ys)
    assert next(script, None) is None, 'the trainer drew fewer shuffles than the reference loop'
    assert trn.batch_size == n // nb
    Eg, Rg = np.asarray(m.E, dtype=np.float64), np.asarray(m.R, dtype=np.float64)

    # --- oracle: the same loop in float64, twice -------------------------------------------------
    # The reference's loop is numerically chaotic: AdaGrad starts from zero accumulators, so the first
    # update of a component is lr * sign(g) however small g is, TransE-L1 gradients are sign() vectors and
    # the margin test is a step function.  A relative perturbation of the initial parameters the size of
    # one fp32 rounding (6e-8) therefore grows to 1e-2 within an epoch IN THE FLOAT64 ORACLE ITSELF.  The
    # "accumulated tolerance" of the fp32 GPU run is measured, not guessed: the GPU may differ from the
    # oracle by no more than a few times what the perturbed oracle run differs from the unperturbed one.
    Eo, Ro = E0.numpy().astype(np.float64), R0.numpy().astype(np.float64)
    viol_cpu = _oracle_fit(orc, kind, Eo, Ro, train.copy(), neg.copy(), perms, nb, margin, lr)
    prng = np.random.default_rng(17)
    Ep = E0.numpy().astype(np.float64) * (1.0 + 6e-8 * prng.standard_normal(E0.shape))
    Rp = R0.numpy().astype(np.float64) * (1.0 + 6e-8 * prng.standard_normal(R0.shape))
    viol_pert = _oracle_fit(orc, kind, Ep, Rp, train.copy(), neg.copy(), perms, nb, margin, lr)

    print(kind, 'violations per epoch: gpu', viol_gpu, 'oracle', viol_cpu, 'perturbed oracle', viol_pert)
    for a, b, c in zip(viol_gpu, viol_cpu, viol_pert):
        assert abs(a - b) <= 4 * abs(c - b) + 0.002 * b, (viol_gpu, viol_cpu, viol_pert)
    for got, want, pert, name in ((Eg, Eo, Ep, 'E'), (Rg, Ro, Rp, 'R')):
        diff, nat = np.abs(got - want), np.abs(pert - want)
        qs = [0.5, 0.9, 0.999]
        dq, nq = np.quantile(diff, qs), np.quantile(nat, qs)
        print('%s %s |gpu - oracle| quantiles %s: %s; |perturbed oracle - oracle|: %s'
              % (kind, name, qs, ['%.2e' % v for v in dq], ['%.2e' % v for v in nq]))
        assert np.all(dq <= 4.0 * nq + 1e-6), (name, dq, nq)

    # --- ranking: on the oracle's final parameters the GPU ranks equal the oracle's ----------------
    m.E.data.copy_(torch.from_numpy(Eo))
    m.R.data.copy_(torch.from_numpy(Ro))
    E32, R32 = np.asarray(m.E, dtype=np.float64), np.asarray(m.R, dtype=np.float64)   # what the GPU holds (fp32)
    Ev = TransEEval if kind == 'transe' else HolEEval
    ev = Ev(test, true)
    pos, fpos = ev.positions(m)
    want_pos, want_fpos, gaps = orc.rank_positions(kind, E32, R32, test, true, tie='count', with_scores=True)
    assert pos == want_pos and fpos == want_fpos
    mrr_o = orc.ranking_scores(want_pos, want_fpos)
    # --- and the two trained models score alike ----------------------------------------------------
    m.E.data.copy_(torch.from_numpy(Eg))
    m.R.data.copy_(torch.from_numpy(Rg))
    pos_g, fpos_g = ev.positions(m)
    mrr_g = orc.ranking_scores(pos_g, fpos_g)
    print(kind, 'oracle-trained (raw, filtered) (mrr, mean rank, hits@10):', mrr_o, 'gpu-trained:', mrr_g)
    for (a, b) in zip(mrr_g, mrr_o):
        assert abs(a[1] - b[1]) <= 0.03 * b[1]              # mean rank
        assert abs(a[2] - b[2]) <= 1.0                      # Hits@10 in percent
