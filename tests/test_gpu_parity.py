"""GPU parity tests: the CUDA path (through the C ABI) against the golden vectors
of the reference and against the CPU oracle on seeded inputs.

Tolerances (BASELINE.json north_star): scores and gradients within 1e-4
relative, parameters within 1e-5 after one step (fp32 arithmetic vs the float64
reference), ranks exactly equal wherever the reference's scores do not tie
within 1e-6.  The oracle is always fed the fp32-rounded parameters cast to
float64, so input rounding is not counted as kernel error.
"""
import numpy as np
import pytest
import torch

from oracle import cpu_oracle as orc

pytestmark = pytest.mark.gpu

GRAD_TOL = dict(rtol=1e-4, atol=2e-6)
PARAM_TOL = dict(rtol=1e-5, atol=2e-6)

PAIRWISE = ['transe_l1_adagrad', 'transe_l2_sgd', 'transe_l1_sgd_d7', 'hole_sigmoid_adagrad',
            'hole_tanh_sgd_rparam', 'hole_linear_adagrad_d256', 'hole_relu_sgd_d10',
            'hole_sigmoid_adagrad_d150']
LOGISTIC = ['hole_logistic_adagrad', 'hole_logistic_sgd_d150', 'rescal_logistic_sgd',
            'rescal_logistic_adagrad_rparam']


def as_xys(tr, ys=None):
    if ys is None:
        return [((int(s), int(o), int(p)), 1.0) for s, o, p in tr]
    return [((int(s), int(o), int(p)), float(y)) for (s, o, p), y in zip(tr, ys)]


def make_model(g):
    import skge
    from skge import activation_functions as afs
    N, d = g['E0'].shape
    if 'W0' in g:
        M = g['W0'].shape[0]
        m = skge.RESCAL((N, N, M), d, rparam=float(g['rparam']))
        m.E[...] = g['E0']
        m.W[...] = g['W0']
        return m
    M = g['R0'].shape[0]
    if 'l1' in g:
        m = skge.TransE((N, N, M), d, l1=bool(g['l1']))
    elif 'af' in g:
        m = skge.HolE((N, N, M), d, rparam=float(g['rparam']), af=afs[str(g['af'])])
    else:
        m = skge.HolE((N, N, M), d, rparam=float(g.get('rparam', 0.0)))
    m.E[...] = g['E0']
    m.R[...] = g['R0']
    return m


def updater_cls(g):
    from skge.param import SGD, AdaGrad
    return AdaGrad if str(g['update']) == 'adagrad' else SGD


def test_device_is_sm100_and_library_loaded():
    from skge import _ext
    assert _ext.lib().skge_check_device() == 0
    assert torch.cuda.get_device_capability()[0] == 10


@pytest.mark.parametrize('name', PAIRWISE)
def test_pairwise_hooks_match_reference_golden(golden, name):
    """_pairwise_gradients + updater calls (the reference's own call sequence)."""
    import skge
    g = golden(name)
    m = make_model(g)
    trn = skge.PairwiseStochasticTrainer(m, nbatches=1, margin=float(g['margin']), max_epochs=1,
                                         learning_rate=float(g['lr']), param_update=updater_cls(g))
    pos, neg = g['pos'], g['neg']
    sc = m._scores(pos[:, 0], pos[:, 2], pos[:, 1])
    np.testing.assert_allclose(sc, g['pscores'] if 'l1' in g else g['raw_p'], rtol=1e-4, atol=1e-5)
    for step in (1, 2):
        grads = m._pairwise_gradients(as_xys(pos), as_xys(neg))
        assert m.nviolations == int(g['nviolations%d' % step])
        assert grads is not None
        np.testing.assert_array_equal(np.asarray(grads['E'][1]), g['eidx%d' % step])
        np.testing.assert_array_equal(np.asarray(grads['R'][1]), g['ridx%d' % step])
        np.testing.assert_allclose(np.asarray(grads['E'][0]), g['ge%d' % step], **GRAD_TOL)
        np.testing.assert_allclose(np.asarray(grads['R'][0]), g['gr%d' % step], **GRAD_TOL)
        trn._batch_step(grads)
        # step 2 inherits step 1's fp32 rounding, hence the slightly wider tolerance there
        tol = PARAM_TOL if step == 1 else dict(rtol=5e-5, atol=1e-5)
        np.testing.assert_allclose(np.asarray(m.E), g['E%d' % step], **tol)
        np.testing.assert_allclose(np.asarray(m.R), g['R%d' % step], **tol)
    if 'violations' in g:
        np.testing.assert_array_equal(m.E.violations, g['violations'])
    if str(g['update']) == 'adagrad':
        uc = np.zeros(g['E0'].shape[0], dtype=int)
        for step in (1, 2):
            uc[g['eidx%d' % step]] += 1
        np.testing.assert_array_equal(m.E.updateCounts, uc)      # skge/param.py:149-150


@pytest.mark.parametrize('name', PAIRWISE)
def test_pairwise_fused_step_matches_reference_golden(golden, name):
    """The fused minibatch kernel sequence (what fit() runs)."""
    import skge
    from skge._modelutil import idx_tensor
    g = golden(name)
    m = make_model(g)
    trn = skge.PairwiseStochasticTrainer(m, nbatches=1, margin=float(g['margin']), max_epochs=1,
                                         learning_rate=float(g['lr']), param_update=updater_cls(g))
    trn._setup_fused()
    pos = tuple(idx_tensor(g['pos'][:, i]) for i in range(3))
    neg = tuple(idx_tensor(g['neg'][:, i]) for i in range(3))
    for step in (1, 2):
        m._fused_pair_step(trn._updaters, pos, neg, None, trn._counts, trn._nviol_dev)
        nv, ue, ur, _ = trn._counts.tolist()
        assert nv == int(g['nviolations%d' % step])
        assert ue == len(g['eidx%d' % step]) and ur == len(g['ridx%d' % step])
        tol = PARAM_TOL if step == 1 else dict(rtol=5e-5, atol=1e-5)
        np.testing.assert_allclose(np.asarray(m.E), g['E%d' % step], **tol)
        np.testing.assert_allclose(np.asarray(m.R), g['R%d' % step], **tol)
    assert int(trn._nviol_dev.item()) == int(g['nviolations1']) + int(g['nviolations2'])


def test_hole_spectra_match_numpy_rfft():
    """Packed spectra (csrc/fft.cuh): slot 0 = (X_0, X_{d/2}), slot f = (Re X_f, Im X_f)."""
    from skge import kernels
    rng = np.random.default_rng(0)
    # powers of two, and mixed radix: d / 2 = 2^a 3^b 5^c (config 2 of BASELINE.json trains d = 150)
    for d in (32, 64, 128, 256, 512, 1024, 36, 48, 50, 54, 60, 90, 96, 100, 150, 160, 200, 250, 300, 486, 750, 1000):
        x = rng.normal(size=(37, d)).astype(np.float32)
        got = kernels.hole_spectra(torch.from_numpy(x).cuda()).cpu().numpy().astype(np.float64)
        F = np.fft.rfft(x.astype(np.float64), axis=1)
        want = np.empty((37, d))
        want[:, 0], want[:, 1] = F[:, 0].real, F[:, d // 2].real
        want[:, 2::2], want[:, 3::2] = F[:, 1:d // 2].real, F[:, 1:d // 2].imag
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5 * np.abs(want).max())


@pytest.mark.parametrize('name', ['hole_linear_adagrad_d256', 'hole_sigmoid_adagrad_d150'])
def test_hole_frequency_domain_step_matches_reference_golden(golden, name):
    """The spectral fused step (what fit() runs for even ncomp with ncomp / 2 = 2^a 3^b 5^c): two steps vs the
    reference's parameters, and the spectra must stay equal to the FFT of the updated tables."""
    import skge
    from skge import kernels
    from skge._modelutil import idx_tensor
    g = golden(name)
    m = make_model(g)
    trn = skge.PairwiseStochasticTrainer(m, nbatches=1, margin=float(g['margin']), max_epochs=1,
                                         learning_rate=float(g['lr']), param_update=updater_cls(g))
    trn._setup_fused()
    m._prepare_fused()
    assert m._spec is not None
    pos = tuple(idx_tensor(g['pos'][:, i]) for i in range(3))
    neg = tuple(idx_tensor(g['neg'][:, i]) for i in range(3))
    for step in (1, 2):
        m._fused_pair_step(trn._updaters, pos, neg, None, trn._counts, trn._nviol_dev)
        nv, ue, ur, _ = trn._counts.tolist()
        assert nv == int(g['nviolations%d' % step])
        assert ue == len(g['eidx%d' % step]) and ur == len(g['ridx%d' % step])
        tol = PARAM_TOL if step == 1 else dict(rtol=5e-5, atol=1e-5)
        np.testing.assert_allclose(np.asarray(m.E), g['E%d' % step], **tol)
        np.testing.assert_allclose(np.asarray(m.R), g['R%d' % step], **tol)
        for tab, hat in ((m.E.data, m._spec[0]), (m.R.data, m._spec[1])):
            ref = kernels.hole_spectra(tab)
            torch.testing.assert_close(hat, ref, rtol=1e-4, atol=1e-5 * float(ref.abs().max()))


@pytest.mark.parametrize('name', LOGISTIC)
def test_logistic_hooks_and_fused_match_reference_golden(golden, name):
    import skge
    from skge._modelutil import idx_tensor
    g = golden(name)
    second = 'W' if 'W0' in g else 'R'
    for fused in (False, True):
        m = make_model(g)
        trn = skge.StochasticTrainer(m, nbatches=1, max_epochs=1, learning_rate=float(g['lr']),
                                     param_update=updater_cls(g))
        xs, ys = g['xs'], g['ys']
        np.testing.assert_allclose(m._scores(xs[:, 0], xs[:, 2], xs[:, 1]), g['scores'], rtol=1e-4, atol=1e-5)
        if fused:
            trn._loss_dev = torch.zeros(1, dtype=torch.float64, device='cuda')
            trn._counts = torch.zeros(4, dtype=torch.int32, device='cuda')
        for step in (1, 2):
            if fused:
                trn._loss_dev.zero_()
                m._fused_logistic_step(trn._updaters, idx_tensor(xs[:, 0]), idx_tensor(xs[:, 1]),
                                       idx_tensor(xs[:, 2]), torch.tensor(ys, dtype=torch.float32, device='cuda'),
                                       trn._counts, trn._loss_dev)
                assert float(trn._loss_dev.item()) == pytest.approx(float(g['loss%d' % step]), rel=1e-4)
                n, ue, u2, _ = trn._counts.tolist()
                assert n == len(xs) and ue == len(g['eidx%d' % step]) and u2 == len(g['idx2_%d' % step])
            else:
                grads = m._gradients(as_xys(xs, ys))
                assert m.loss == pytest.approx(float(g['loss%d' % step]), rel=1e-4)
                np.testing.assert_array_equal(np.asarray(grads['E'][1]), g['eidx%d' % step])
                np.testing.assert_array_equal(np.asarray(grads[second][1]), g['idx2_%d' % step])
                np.testing.assert_allclose(np.asarray(grads['E'][0]), g['ge%d' % step], **GRAD_TOL)
                np.testing.assert_allclose(np.asarray(grads[second][0]), g['g2_%d' % step], **GRAD_TOL)
                trn._batch_step(grads)
            tol = PARAM_TOL if step == 1 else dict(rtol=5e-5, atol=1e-5)
            np.testing.assert_allclose(np.asarray(m.E), g['E%d' % step], **tol)
            np.testing.assert_allclose(np.asarray(getattr(m, second)), g['P2_%d' % step], **tol)


@pytest.mark.parametrize('name', ['hole_logistic_sgd_d150', 'rescal_logistic_sgd'])
def test_logistic_step_skips_masked_examples(golden, name):
    """Examples the sampler could not produce are masked, not compacted: the step must equal
    the oracle's step on the valid subset (skge/sample.py:22-24, skge/base.py:1295-1304)."""
    import skge
    from skge._modelutil import idx_tensor
    g = golden(name)
    rescal = 'W0' in g
    xs, ys = g['xs'], g['ys']
    rng = np.random.default_rng(1)
    keep = rng.random(len(xs)) < 0.7
    m = make_model(g)
    trn = skge.StochasticTrainer(m, nbatches=1, max_epochs=1, learning_rate=float(g['lr']), param_update=updater_cls(g))
    trn._loss_dev = torch.zeros(1, dtype=torch.float64, device='cuda')
    trn._counts = torch.zeros(4, dtype=torch.int32, device='cuda')
    m._fused_logistic_step(trn._updaters, idx_tensor(xs[:, 0]), idx_tensor(xs[:, 1]), idx_tensor(xs[:, 2]),
                           torch.tensor(ys, dtype=torch.float32, device='cuda'), trn._counts, trn._loss_dev,
                           valid=torch.tensor(keep.astype(np.uint8), device='cuda'))
    E, P2 = g['E0'].copy(), (g['W0'] if rescal else g['R0']).copy()
    fn = orc.rescal_gradients if rescal else orc.hole_gradients
    grads, loss = fn(E, P2, xs[keep], ys[keep], float(g['rparam']))
    k2 = 'W' if rescal else 'R'
    orc.sgd_update(E, grads['E'][0], grads['E'][1], float(g['lr']), None if rescal else 'normless1')
    orc.sgd_update(P2, grads[k2][0], grads[k2][1], float(g['lr']), None)
    n, ue, u2, _ = trn._counts.tolist()
    assert n == int(keep.sum()) and ue == len(grads['E'][1]) and u2 == len(grads[k2][1])
    assert float(trn._loss_dev.item()) == pytest.approx(loss, rel=1e-4)
    np.testing.assert_allclose(np.asarray(m.E), E, **PARAM_TOL)
    np.testing.assert_allclose(np.asarray(getattr(m, k2)), P2, **PARAM_TOL)


def test_appendix_a_known_answers(golden):
    import skge
    g = golden('appendix_a')
    E0, R0, pos, neg = g['E0'], g['R0'], g['pos'], g['neg']
    m = skge.TransE((5, 5, 2), 4, l1=True)
    m.E[...] = E0
    m.R[...] = R0
    trn = skge.PairwiseStochasticTrainer(m, nbatches=1, margin=2.0, learning_rate=0.1)
    grads = m._pairwise_gradients(as_xys(pos), as_xys(neg))
    assert m.nviolations == 4
    np.testing.assert_allclose(np.asarray(grads['E'][0]), g['a1_ge'], **GRAD_TOL)
    np.testing.assert_allclose(np.asarray(grads['R'][0]), g['a1_gr'], **GRAD_TOL)
    trn._batch_step(grads)
    np.testing.assert_allclose(np.asarray(m.E), g['a1_E'], **PARAM_TOL)
    np.testing.assert_allclose(np.asarray(m.R), g['a1_R'], **PARAM_TOL)
    m = skge.HolE((5, 5, 2), 4)
    m.E[...] = E0
    m.R[...] = R0
    trn = skge.PairwiseStochasticTrainer(m, nbatches=1, margin=0.2, learning_rate=0.1)
    np.testing.assert_allclose(m._scores(pos[:, 0], pos[:, 2], pos[:, 1]), g['a2_raw_p'], rtol=1e-5, atol=1e-6)
    grads = m._pairwise_gradients(as_xys(pos), as_xys(neg))
    np.testing.assert_allclose(np.asarray(grads['E'][0]), g['a2_ge'], **GRAD_TOL)
    np.testing.assert_allclose(np.asarray(grads['R'][0]), g['a2_gr'], **GRAD_TOL)
    trn._batch_step(grads)
    np.testing.assert_allclose(np.asarray(m.E), g['a2_E'], **PARAM_TOL)
    np.testing.assert_allclose(np.asarray(m.R), g['a2_R'], **PARAM_TOL)
    m = skge.RESCAL((5, 5, 2), 4)
    m.E[...] = E0
    m.W[...] = g['W0']
    grads = m._gradients(as_xys(g['a3_xs'], g['a3_ys']))
    assert m.loss == pytest.approx(4.651044182920225, rel=1e-5)
    np.testing.assert_allclose(np.asarray(grads['E'][0]), g['a3_ge'], **GRAD_TOL)
    np.testing.assert_allclose(np.asarray(grads['W'][0]), g['a3_gw'], **GRAD_TOL)
    np.testing.assert_array_equal(np.asarray(grads['W'][1]), g['a3_pidx'])


def test_no_violation_returns_none():
    """skge/transe.py:90-91, skge/hole.py:58-59."""
    import skge
    for cls in (skge.TransE, skge.HolE):
        m = cls((6, 6, 2), 8)
        skge.PairwiseStochasticTrainer(m, margin=-1e6)
        pos = [((0, 1, 0), 1.0), ((2, 3, 1), 1.0)]
        neg = [((4, 1, 0), 1.0), ((2, 5, 1), 1.0)]
        assert m._pairwise_gradients(pos, neg) is None
        assert m.nviolations == 0


def _full_size_batch(kind, N, M, d, B, seed):
    """One minibatch of a WN18-shaped problem: random parameters after the
    model's own post-hook, B positives x 2 corruptions."""
    rng = np.random.default_rng(seed)
    E = rng.uniform(-1, 1, (N, d)) * (6.0 / np.sqrt(N + d)) * 20
    R = rng.uniform(-1, 1, (M, d)) * 0.3
    E = orc.normalize(E) if kind == 'transe' else E * 0.2
    E, R = E.astype(np.float32).astype(np.float64), R.astype(np.float32).astype(np.float64)
    pos = np.stack([rng.integers(N, size=B), rng.integers(N, size=B), rng.integers(M, size=B)], 1)
    pos = np.repeat(pos, 2, axis=0)
    neg = pos.copy()
    neg[0::2, 0] = rng.integers(N, size=B)
    neg[1::2, 1] = rng.integers(N, size=B)
    return E, R, pos, neg


def _near_margin(info, margin, tol):
    return np.abs(info['nscores'] + margin - info['pscores']) < tol


@pytest.mark.parametrize('kind,d,margin,l1,opt', [
    ('transe', 50, 2.0, True, 'adagrad'), ('transe', 200, 2.0, True, 'adagrad'), ('transe', 64, 1.0, False, 'adagrad'),
    ('hole', 150, 0.2, None, 'adagrad'),
    # power-of-two d: the shared-memory FFT path (SGD: the update is continuous in g, so the
    # whole table can be compared without excusing AdaGrad's clamp region)
    ('hole', 256, 0.2, None, 'sgd'), ('hole', 128, 0.2, None, 'sgd'), ('hole', 32, 0.2, None, 'sgd'),
    ('hole', 1024, 0.2, None, 'sgd'), ('hole', 256, 0.2, None, 'adagrad'),
    # mixed-radix row lengths (d / 2 = 2^a 3^b 5^c): 150 is config 2; 200 and 160 have the access shape of
    # d = 256 (two 128-bit chunks per lane) and must NOT take its register-resident transform
    ('hole', 150, 0.2, None, 'sgd'), ('hole', 200, 0.2, None, 'sgd'), ('hole', 100, 0.2, None, 'sgd'),
    ('hole', 96, 0.2, None, 'sgd'), ('hole', 160, 0.2, None, 'sgd'), ('hole', 250, 0.2, None, 'sgd'),
    ('hole', 54, 0.2, None, 'sgd'), ('hole', 600, 0.2, None, 'sgd')])
def test_wn18_shaped_minibatch_against_oracle(kind, d, margin, l1, opt):
    """Config-1/2/4-sized minibatch (B = 1414 -> P = 2828 pairs) vs the oracle."""
    import skge
    from skge.param import AdaGrad, SGD
    N, M, B = 40943, 18, 1414
    E0, R0, pos, neg = _full_size_batch(kind, N, M, d, B, seed=d)
    if kind == 'transe':
        m = skge.TransE((N, N, M), d, l1=l1)
        ograds, info = orc.transe_pairwise_gradients(E0, R0, pos, neg, margin, l1)
    else:
        m = skge.HolE((N, N, M), d)
        ograds, info = orc.hole_pairwise_gradients(E0, R0, pos, neg, margin, 'sigmoid', 0.0)
    m.E[...] = E0
    m.R[...] = R0
    trn = skge.PairwiseStochasticTrainer(m, nbatches=1, margin=margin, learning_rate=0.1,
                                         param_update=AdaGrad if opt == 'adagrad' else SGD)
    grads = m._pairwise_gradients(as_xys(pos), as_xys(neg))
    # pairs within 1e-5 of the margin may flip between fp32 and float64
    assert not _near_margin(info, margin, 1e-5).any(), 'regenerate the case: a pair sits on the margin'
    assert m.nviolations == info['nviolations']
    np.testing.assert_array_equal(np.asarray(grads['E'][1]), ograds['E'][1])
    np.testing.assert_array_equal(np.asarray(grads['R'][1]), ograds['R'][1])
    np.testing.assert_allclose(np.asarray(grads['E'][0]), ograds['E'][0], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(np.asarray(grads['R'][0]), ograds['R'][0], rtol=1e-4, atol=1e-5)
    trn._batch_step(grads)
    E, R = E0.copy(), R0.copy()
    post = 'normalize' if kind == 'transe' else 'normless1'
    if opt == 'sgd':
        orc.sgd_update(E, ograds['E'][0], ograds['E'][1], 0.1, post)
        orc.sgd_update(R, ograds['R'][0], ograds['R'][1], 0.1, None)
        np.testing.assert_allclose(np.asarray(m.E), E, **PARAM_TOL)
        np.testing.assert_allclose(np.asarray(m.R), R, **PARAM_TOL)
        if kind == 'hole':      # the same minibatch through the frequency-domain fused step
            from skge._modelutil import idx_tensor
            m2 = skge.HolE((N, N, M), d)
            m2.E[...] = E0
            m2.R[...] = R0
            t2 = skge.PairwiseStochasticTrainer(m2, nbatches=1, margin=margin, learning_rate=0.1, param_update=SGD)
            t2._setup_fused()
            m2._prepare_fused()
            assert m2._spec is not None
            m2._fused_pair_step(t2._updaters, tuple(idx_tensor(pos[:, i]) for i in range(3)),
                                tuple(idx_tensor(neg[:, i]) for i in range(3)), None, t2._counts, t2._nviol_dev)
            assert int(t2._nviol_dev.item()) == info['nviolations']
            np.testing.assert_allclose(np.asarray(m2.E), E, **PARAM_TOL)
            np.testing.assert_allclose(np.asarray(m2.R), R, **PARAM_TOL)
        return
    orc.adagrad_update(E, np.zeros_like(E), ograds['E'][0], ograds['E'][1], 0.1, post)
    orc.adagrad_update(R, np.zeros_like(R), ograds['R'][0], ograds['R'][1], 0.1, None)
    # AdaGrad's first step is x -= lr * g / max(|g|, 1e-7) (skge/param.py:147-155): for |g| below
    # ~1e-6 the step is a discontinuous function of g (sign flip / clamp region), so fp32 and
    # float64 legitimately disagree there.  Such components are excluded -- and must be rare.
    # The row post-hook spreads such a component's difference over its whole row, so the rows
    # that contain one are excluded.
    gotE, ok = np.asarray(m.E, dtype=np.float64), np.ones(E.shape[0], dtype=bool)
    ok[ograds['E'][1]] = ((np.abs(ograds['E'][0]) > 1e-6) | (ograds['E'][0] == 0)).all(axis=1)
    assert (~ok).mean() < 5e-3
    np.testing.assert_allclose(gotE[ok], E[ok], **PARAM_TOL)
    np.testing.assert_allclose(np.asarray(m.R), R, **PARAM_TOL)


@pytest.mark.parametrize('kind,opt', [('hole', 'sgd'), ('hole', 'adagrad'), ('transe', 'sgd'), ('transe', 'adagrad')])
def test_large_minibatch_staged_update_against_oracle(kind, opt):
    """More than 2^18 occurrences of 1 KB rows (d = 256): the bulk-TMA staged segmented update
    (csrc/segment.cu::seg_reduce_bulk_kernel), for HolE with the register-resident transforms and
    the relation rows pre-reduced over relation-ordered pairs, a hub entity and a frequent
    relation on the chunked path, pairs of every sharing pattern mixed in."""
    import skge
    from skge.param import AdaGrad, SGD
    from skge._modelutil import idx_tensor
    N, M, B, d = 30000, 40, 23000, 256
    E0, R0, pos, neg = _full_size_batch(kind, N, M, d, B, seed=7 + (kind == 'hole'))
    rng = np.random.default_rng(3)
    P = len(pos)
    assert 6 * P > (1 << 18)
    hub = rng.choice(P, 3000, replace=False)
    pos[hub, 0] = 17                                  # hub subject
    neg[hub, 0] = np.where(neg[hub, 1] != pos[hub, 1], 17, neg[hub, 0])   # keep each pair's corruption pattern
    pos[hub[:1500], 2] = 5                            # frequent relation
    neg[hub[:1500], 2] = 5
    odd = rng.choice(P, 600, replace=False)           # other shapes: predicate / both entities corrupted, neg == pos
    neg[odd[:200], 2] = rng.integers(M, size=200)
    neg[odd[200:400], 0] = rng.integers(N, size=200)
    neg[odd[200:400], 1] = rng.integers(N, size=200)
    neg[odd[400:]] = pos[odd[400:]]
    margin = 2.0 if kind == 'transe' else 0.2
    if kind == 'transe':
        m = skge.TransE((N, N, M), d, l1=True)
        ograds, info = orc.transe_pairwise_gradients(E0, R0, pos, neg, margin, True)
    else:
        m = skge.HolE((N, N, M), d)
        ograds, info = orc.hole_pairwise_gradients(E0, R0, pos, neg, margin, 'sigmoid', 0.0)
    near = _near_margin(info, margin, 1e-5)
    keep = ~near                                      # pairs on the margin may flip between fp32 and float64
    if near.any():
        pos, neg = pos[keep], neg[keep]
        assert 6 * len(pos) > (1 << 18)
        if kind == 'transe':
            ograds, info = orc.transe_pairwise_gradients(E0, R0, pos, neg, margin, True)
        else:
            ograds, info = orc.hole_pairwise_gradients(E0, R0, pos, neg, margin, 'sigmoid', 0.0)
    m.E[...] = E0
    m.R[...] = R0
    trn = skge.PairwiseStochasticTrainer(m, nbatches=1, margin=margin, learning_rate=0.1,
                                         param_update=AdaGrad if opt == 'adagrad' else SGD)
    trn._setup_fused()
    if kind == 'hole':
        m._prepare_fused()
        assert m._spec is not None
    m._fused_pair_step(trn._updaters, tuple(idx_tensor(pos[:, i]) for i in range(3)),
                       tuple(idx_tensor(neg[:, i]) for i in range(3)), None, trn._counts, trn._nviol_dev)
    nv, ue, ur, _ = trn._counts.tolist()
    assert nv == info['nviolations'] and ue == len(ograds['E'][1]) and ur == len(ograds['R'][1])
    E, R = E0.copy(), R0.copy()
    post = 'normalize' if kind == 'transe' else 'normless1'
    if opt == 'sgd':
        orc.sgd_update(E, ograds['E'][0], ograds['E'][1], 0.1, post)
        orc.sgd_update(R, ograds['R'][0], ograds['R'][1], 0.1, None)
        np.testing.assert_allclose(np.asarray(m.E), E, **PARAM_TOL)
        np.testing.assert_allclose(np.asarray(m.R), R, **PARAM_TOL)
    else:
        orc.adagrad_update(E, np.zeros_like(E), ograds['E'][0], ograds['E'][1], 0.1, post)
        orc.adagrad_update(R, np.zeros_like(R), ograds['R'][0], ograds['R'][1], 0.1, None)
        # AdaGrad's first step is discontinuous in g near 0 (see test_wn18_shaped_minibatch_against_oracle)
        gotE, ok = np.asarray(m.E, dtype=np.float64), np.ones(E.shape[0], dtype=bool)
        ok[ograds['E'][1]] = ((np.abs(ograds['E'][0]) > 1e-6) | (ograds['E'][0] == 0)).all(axis=1)
        assert (~ok).mean() < 2e-2
        np.testing.assert_allclose(gotE[ok], E[ok], **PARAM_TOL)
        okr = np.ones(R.shape[0], dtype=bool)
        okr[ograds['R'][1]] = ((np.abs(ograds['R'][0]) > 1e-6) | (ograds['R'][0] == 0)).all(axis=1)
        np.testing.assert_allclose(np.asarray(m.R, dtype=np.float64)[okr], R[okr], **PARAM_TOL)
    if kind == 'hole':       # the spectra the step maintains must be the transforms of the updated tables
        from skge import kernels
        for tab, hat in ((m.E.data, m._spec[0]), (m.R.data, m._spec[1])):
            ref = kernels.hole_spectra(tab)
            torch.testing.assert_close(hat, ref, rtol=1e-4, atol=1e-5 * float(ref.abs().max()))


@pytest.mark.parametrize('kind,d', [('transe', 50), ('transe', 256), ('hole', 64), ('hole', 150)])
def test_hot_rows_use_the_chunked_segment_reduction(kind, d):
    """Rows with thousands of occurrences in one minibatch (a hub entity, a frequent
    relation) go through the chunked two-level reduction; the mean must still match."""
    import skge
    from skge.param import SGD
    N, M, P = 40, 2, 6000
    rng = np.random.default_rng(d + 1)
    E0 = (rng.normal(size=(N, d)) * 0.3).astype(np.float32).astype(np.float64)
    R0 = (rng.normal(size=(M, d)) * 0.3).astype(np.float32).astype(np.float64)
    if kind == 'transe':
        E0 = orc.normalize(E0).astype(np.float32).astype(np.float64)
    pos = np.stack([rng.integers(N, size=P), rng.integers(N, size=P), rng.integers(M, size=P)], 1)
    pos[: P // 2, 0] = 7                      # hub subject: > 3000 occurrences
    neg = pos.copy()
    neg[0::2, 0] = rng.integers(N, size=P // 2)
    neg[1::2, 1] = rng.integers(N, size=P // 2)
    margin = 2.0 if kind == 'transe' else 0.2
    if kind == 'transe':
        m = skge.TransE((N, N, M), d)
        ograds, info = orc.transe_pairwise_gradients(E0, R0, pos, neg, margin, True)
    else:
        m = skge.HolE((N, N, M), d, rparam=0.05)
        ograds, info = orc.hole_pairwise_gradients(E0, R0, pos, neg, margin, 'sigmoid', 0.05)
    m.E[...] = E0
    m.R[...] = R0
    trn = skge.PairwiseStochasticTrainer(m, nbatches=1, margin=margin, learning_rate=0.1, param_update=SGD)
    assert not _near_margin(info, margin, 1e-5).any()
    assert np.bincount(np.concatenate([pos[info['ind'], 0], neg[info['ind'], 0]])).max() > 1000
    grads = m._pairwise_gradients(as_xys(pos), as_xys(neg))
    assert m.nviolations == info['nviolations']
    np.testing.assert_array_equal(np.asarray(grads['E'][1]), ograds['E'][1])
    np.testing.assert_array_equal(np.asarray(grads['R'][1]), ograds['R'][1])
    np.testing.assert_allclose(np.asarray(grads['E'][0]), ograds['E'][0], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(np.asarray(grads['R'][0]), ograds['R'][0], rtol=1e-4, atol=1e-5)
    # and the fused step (SGD: continuous in g)
    trn._setup_fused()
    from skge._modelutil import idx_tensor
    m._fused_pair_step(trn._updaters, tuple(idx_tensor(pos[:, i]) for i in range(3)),
                       tuple(idx_tensor(neg[:, i]) for i in range(3)), None, trn._counts, trn._nviol_dev)
    E, R = E0.copy(), R0.copy()
    orc.sgd_update(E, ograds['E'][0], ograds['E'][1], 0.1, 'normalize' if kind == 'transe' else 'normless1')
    orc.sgd_update(R, ograds['R'][0], ograds['R'][1], 0.1, None)
    np.testing.assert_allclose(np.asarray(m.E), E, **PARAM_TOL)
    np.testing.assert_allclose(np.asarray(m.R), R, **PARAM_TOL)
    if kind == 'hole':
        # the same minibatch through the frequency-domain step: hub rows finish in
        # seg_long_finish_kernel's spectral branch (mixed-radix transform for d = 150)
        from skge import kernels
        m2 = skge.HolE((N, N, M), d, rparam=0.05)
        m2.E[...] = E0
        m2.R[...] = R0
        t2 = skge.PairwiseStochasticTrainer(m2, nbatches=1, margin=margin, learning_rate=0.1, param_update=SGD)
        t2._setup_fused()
        m2._prepare_fused()
        assert m2._spec is not None
        m2._fused_pair_step(t2._updaters, tuple(idx_tensor(pos[:, i]) for i in range(3)),
                            tuple(idx_tensor(neg[:, i]) for i in range(3)), None, t2._counts, t2._nviol_dev)
        np.testing.assert_allclose(np.asarray(m2.E), E, **PARAM_TOL)
        np.testing.assert_allclose(np.asarray(m2.R), R, **PARAM_TOL)
        for tab, hat in ((m2.E.data, m2._spec[0]), (m2.R.data, m2._spec[1])):
            ref = kernels.hole_spectra(tab)
            torch.testing.assert_close(hat, ref, rtol=1e-4, atol=1e-5 * float(ref.abs().max()))


@pytest.mark.parametrize('N,M,d,B', [
    (40943, 18, 100, 1414),     # config 3: d = 100, 1414 positives + 2828 negatives
    (3000, 700, 36, 500),       # many relations: several runs of W[p] inside one 32-example chunk
    (2000, 5, 160, 300),        # 128 < d: the 256-column layout (16 examples per chunk)
    (600, 3, 7, 100),           # tiny odd row length (padded to 8 in shared memory)
    (500, 3, 230, 100)])        # W[p] too large for shared memory: one CTA per example
def test_rescal_wn18_shaped_minibatch_against_oracle(N, M, d, B):
    """Logistic loss, SGD, relation-grouped kernel (csrc/logistic.cu::rescal_logistic_grouped_kernel)."""
    import skge
    from skge.param import SGD
    rng = np.random.default_rng(3)
    E0 = (rng.uniform(-1, 1, (N, d)) * 0.1).astype(np.float32).astype(np.float64)
    W0 = (rng.uniform(-1, 1, (M, d, d)) * 0.1).astype(np.float32).astype(np.float64)
    xs = np.stack([rng.integers(N, size=3 * B), rng.integers(N, size=3 * B), rng.integers(M, size=3 * B)], 1)
    ys = np.concatenate([np.ones(B), -np.ones(2 * B)])
    m = skge.RESCAL((N, N, M), d, rparam=0.01)
    m.E[...] = E0
    m.W[...] = W0
    trn = skge.StochasticTrainer(m, nbatches=1, learning_rate=0.1, param_update=SGD)
    ograds, oloss = orc.rescal_gradients(E0, W0, xs, ys, 0.01)
    grads = m._gradients(as_xys(xs, ys))
    assert m.loss == pytest.approx(oloss, rel=1e-5)
    np.testing.assert_array_equal(np.asarray(grads['E'][1]), ograds['E'][1])
    np.testing.assert_array_equal(np.asarray(grads['W'][1]), ograds['W'][1])
    np.testing.assert_allclose(np.asarray(grads['E'][0]), ograds['E'][0], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(np.asarray(grads['W'][0]), ograds['W'][0], rtol=1e-4, atol=1e-6)
    trn._batch_step(grads)
    E, W = E0.copy(), W0.copy()
    orc.sgd_update(E, ograds['E'][0], ograds['E'][1], 0.1)
    orc.sgd_update(W, ograds['W'][0], ograds['W'][1], 0.1)
    np.testing.assert_allclose(np.asarray(m.E), E, **PARAM_TOL)
    np.testing.assert_allclose(np.asarray(m.W), W, **PARAM_TOL)


@pytest.mark.parametrize('post', [None, 'normalize', 'normless1'])
@pytest.mark.parametrize('opt', ['sgd', 'adagrad'])
@pytest.mark.parametrize('d', [7, 50, 150, 256, 1000])
def test_sparse_update_and_post_hooks(opt, post, d):
    """ParameterUpdate.__call__(g, idx) (skge/param.py:115-118) incl. empty idx."""
    from skge.param import Parameter, SGD, AdaGrad, normalize, normless1
    rng = np.random.default_rng(d)
    X0 = (rng.normal(size=(300, d)) * (2.0 if post == 'normless1' else 0.3)).astype(np.float32).astype(np.float64)
    hook = {None: None, 'normalize': normalize, 'normless1': normless1}[post]
    p = Parameter(X0.shape, 'nunif', name='X', post=hook, value=X0)
    upd = (SGD if opt == 'sgd' else AdaGrad)(p, 0.1)
    X, p2 = X0.copy(), np.zeros_like(X0)
    for it in range(2):
        idx = np.unique(rng.integers(300, size=120))
        g = rng.normal(size=(len(idx), d)).astype(np.float32).astype(np.float64)
        upd(g, idx)
        if opt == 'sgd':
            orc.sgd_update(X, g, idx, 0.1, post)
        else:
            orc.adagrad_update(X, p2, g, idx, 0.1, post)
        np.testing.assert_allclose(np.asarray(p), X, **(PARAM_TOL if it == 0 else dict(rtol=5e-5, atol=1e-5)))
    upd(np.zeros((0, d)), np.zeros(0, dtype=np.int64))          # empty update is a no-op
    np.testing.assert_allclose(np.asarray(p), X, rtol=5e-5, atol=1e-5)


def test_parameter_init_and_quirks():
    """init ranges (skge/param.py:11-54), post at creation (:72-73), the
    column-wise normless1 quirk (:171-173), error behaviour (:100-103)."""
    from skge.param import Parameter, normalize, normless1
    p = Parameter((1000, 50), 'nunif', name='E', post=normalize)
    a = np.asarray(p, dtype=np.float64)
    np.testing.assert_allclose(np.linalg.norm(a, axis=1), 1.0, rtol=1e-5)
    p = Parameter((1000, 50), 'nunif')
    a = np.asarray(p)
    bnd = np.sqrt(6) / np.sqrt(1050)
    assert np.abs(a).max() <= bnd * (1 + 1e-6) and a.std() == pytest.approx(bnd / np.sqrt(3), rel=0.05)
    a = np.asarray(Parameter((400, 30), 'unif'))
    assert np.abs(a).max() <= 1 / np.sqrt(400) * (1 + 1e-6)
    w = Parameter((3, 20, 20), 'nunif')
    assert w.shape == (3, 20, 20)
    raw = Parameter((2000, 16), 'randn')
    q = Parameter((2000, 16), 'randn', post=normless1, value=None)
    colsq = (np.asarray(q, dtype=np.float64) ** 2).sum(axis=0)
    assert (colsq < 1.0 / 1500).all()       # columns were divided by ~N (sum of N unit-variance squares)
    assert raw.shape == q.shape
    with pytest.raises(ValueError, match='Unknown initialization'):
        Parameter((4, 4), 'bogus')
    with pytest.raises(ValueError, match='Shape must be of size 2'):
        Parameter((4,), 'nunif')


def test_device_sampler_invariants(golden):
    """RandomModeSampler contract (skge/sample.py:17-46) on the device."""
    from skge.sample import RandomModeSampler, LCWASampler
    g = golden('sampler')
    xs = [tuple(t) for t in g['xs'].tolist()]
    N, M = int(g['N']), int(g['M'])
    smp = RandomModeSampler(1, [0, 1], xs, (N, N, M))
    res = smp.sample([(x, 1.0) for x in xs])
    assert len(res) == int(g['nneg']) == 2 * len(xs)       # same yield as the reference on this graph
    xs_set = set(xs)
    for i, (nx, y) in enumerate(res):
        x = xs[i // 2]
        assert y == -1.0 and nx not in xs_set and nx[2] == x[2]
        if i % 2 == 0:
            assert nx[1] == x[1] and 0 <= nx[0] < N        # subject corrupted first
        else:
            assert nx[0] == x[0] and 0 <= nx[1] < N
    # a saturated slot is skipped after ntries (sample.py:22-24): all (s, 0, 0) exist
    full = [(s, 0, 0) for s in range(N)]
    smp = RandomModeSampler(1, [0], full, (N, N, 1))
    assert smp.sample([(full[0], 1.0)]) == []
    # draws are uniform over the admissible values
    big = RandomModeSampler(200, [0], [(0, 0, 0)], (8, 8, 1))
    draws = np.array([nx[0] for nx, _ in big.sample([((0, 0, 0), 1.0)] * 50)])
    cnt = np.bincount(draws, minlength=8)
    assert cnt[0] == 0 and cnt[1:].min() > 0.8 * cnt[1:].mean()
    # LCWA: the corrupted triple keeps an (s, p) that occurs in training (sample.py:103-110)
    lc = LCWASampler(2, [0, 1, 2], xs, (N, N, M))
    seen_sp = {(s, p) for s, o, p in xs}
    for nx, _ in lc.sample([(x, 1.0) for x in xs]):
        assert nx not in xs_set and (nx[0], nx[2]) in seen_sp


def test_tripleset_membership():
    from skge import kernels
    from skge._modelutil import idx_tensor
    rng = np.random.default_rng(0)
    tr = np.unique(np.stack([rng.integers(5000, size=20000), rng.integers(5000, size=20000),
                             rng.integers(40, size=20000)], 1), axis=0)
    ts = kernels.TripleSet(*(idx_tensor(tr[:, i]) for i in range(3)), 5000, 40)
    assert ts.contains(*(idx_tensor(tr[:, i]) for i in range(3))).all()
    other = tr.copy()
    other[:, 1] = (other[:, 1] + 1 + rng.integers(4000, size=len(tr))) % 5000
    truth = np.array([tuple(t) in set(map(tuple, tr.tolist())) for t in other.tolist()])
    got = ts.contains(*(idx_tensor(other[:, i]) for i in range(3))).cpu().numpy().astype(bool)
    np.testing.assert_array_equal(got, truth)


@pytest.mark.parametrize('kind,d', [('transe', 200), ('hole', 256), ('hole', 150)])
def test_fused_step_is_bit_reproducible(kind, d):
    """No atomics touch a parameter row and hot rows are reduced in a fixed order, so the same
    minibatch from the same state must give the same bits (the race check this pool's closed
    compute-sanitizer cannot run)."""
    import skge
    from skge._modelutil import idx_tensor
    rng = np.random.default_rng(d)
    N, M, P = 3000, 5, 20000                      # ~27 occurrences per entity, 8000 per relation
    pos = np.stack([rng.integers(N, size=P), rng.integers(N, size=P), rng.integers(M, size=P)], 1)
    pos[: P // 4, 1] = 11                         # a hub entity
    neg = pos.copy()
    neg[0::2, 0] = rng.integers(N, size=P // 2)
    neg[1::2, 1] = rng.integers(N, size=P // 2)
    E0 = (rng.normal(size=(N, d)) * 0.3).astype(np.float32)
    R0 = (rng.normal(size=(M, d)) * 0.3).astype(np.float32)
    outs = []
    for rep in range(3):
        m = (skge.TransE if kind == 'transe' else skge.HolE)((N, N, M), d)
        m.E[...] = E0
        m.R[...] = R0
        trn = skge.PairwiseStochasticTrainer(m, nbatches=1, margin=2.0 if kind == 'transe' else 0.2, learning_rate=0.1)
        trn._setup_fused()
        if hasattr(m, '_prepare_fused'):
            m._prepare_fused()
        for _ in range(2):
            m._fused_pair_step(trn._updaters, tuple(idx_tensor(pos[:, i]) for i in range(3)),
                               tuple(idx_tensor(neg[:, i]) for i in range(3)), None, trn._counts, trn._nviol_dev)
        outs.append((m.E.data.clone(), m.R.data.clone(), trn._updaters['E'].p2.clone()))
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            assert torch.equal(a, b)


@pytest.mark.parametrize('d', [150, 64, 24, 128, 256])
def test_hole_twin_rows_fold_for_every_sharing_pattern(d):
    """The HolE kernels sum the positive's and the negative's contribution to a shared slot
    into one gradient row that counts twice in the mean.  Negatives that share 0, 1, 2 or all 3
    slots with their positive (supplied-negatives mode, predicate corruption, a pair scored
    against itself) must all give the reference's means: direct (d = 150, 24), shared-memory FFT
    (d = 64, 128, 256) and, for those, the frequency-domain fused step (d = 128 / 256: staged pairs
    of the two corrupted shapes, every other shape through the six-row path)."""
    import skge
    from skge.param import SGD
    from skge._modelutil import idx_tensor
    N, M, B = 500, 7, 600
    rng = np.random.default_rng(d)
    E0 = (rng.uniform(-1, 1, (N, d)) * 0.2).astype(np.float32).astype(np.float64)
    R0 = (rng.uniform(-1, 1, (M, d)) * 0.3).astype(np.float32).astype(np.float64)
    pos = np.stack([rng.integers(N, size=B), rng.integers(N, size=B), rng.integers(M, size=B)], 1)
    neg = pos.copy()
    kind = np.arange(B) % 6
    neg[kind == 0, 0] = rng.integers(N, size=(kind == 0).sum())          # subject corrupted
    neg[kind == 1, 1] = rng.integers(N, size=(kind == 1).sum())          # object corrupted
    neg[kind == 2, 2] = rng.integers(M, size=(kind == 2).sum())          # predicate corrupted
    k3 = kind == 3                                                       # nothing shared
    neg[k3] = np.stack([rng.integers(N, size=k3.sum()), rng.integers(N, size=k3.sum()),
                        rng.integers(M, size=k3.sum())], 1)
    k4 = kind == 4                                                       # subject and object corrupted
    neg[k4, 0] = rng.integers(N, size=k4.sum())
    neg[k4, 1] = rng.integers(N, size=k4.sum())
    # kind == 5: the negative IS the positive (violates for every margin > 0)
    pos[::50, 1] = pos[::50, 0]                                          # self loops: s == o
    neg[::50] = pos[::50]
    neg[::50, 0] = (pos[::50, 0] + 1) % N
    margin = 0.2
    ograds, info = orc.hole_pairwise_gradients(E0, R0, pos, neg, margin, 'sigmoid', 0.0)
    assert not _near_margin(info, margin, 1e-5).any()
    m = skge.HolE((N, N, M), d)
    m.E[...] = E0
    m.R[...] = R0
    skge.PairwiseStochasticTrainer(m, nbatches=1, margin=margin, learning_rate=0.1, param_update=SGD)  # sets m.margin
    grads = m._pairwise_gradients(as_xys(pos), as_xys(neg))
    assert m.nviolations == info['nviolations']
    for k in ('E', 'R'):
        np.testing.assert_array_equal(np.asarray(grads[k][1]), ograds[k][1])
        np.testing.assert_allclose(np.asarray(grads[k][0]), ograds[k][0], rtol=1e-4, atol=1e-6)
    E, R = E0.copy(), R0.copy()
    orc.sgd_update(E, ograds['E'][0], ograds['E'][1], 0.1, 'normless1')
    orc.sgd_update(R, ograds['R'][0], ograds['R'][1], 0.1, None)
    m2 = skge.HolE((N, N, M), d)
    m2.E[...] = E0
    m2.R[...] = R0
    t2 = skge.PairwiseStochasticTrainer(m2, nbatches=1, margin=margin, learning_rate=0.1, param_update=SGD)
    t2._setup_fused()
    m2._prepare_fused()
    assert (m2._spec is not None) == (d != 24)      # 150 = 2 * 3 * 5 * 5: mixed-radix transform; 24 < 32: direct kernel
    m2._fused_pair_step(t2._updaters, tuple(idx_tensor(pos[:, i]) for i in range(3)),
                        tuple(idx_tensor(neg[:, i]) for i in range(3)), None, t2._counts, t2._nviol_dev)
    assert int(t2._nviol_dev.item()) == info['nviolations']
    np.testing.assert_allclose(np.asarray(m2.E), E, **PARAM_TOL)
    np.testing.assert_allclose(np.asarray(m2.R), R, **PARAM_TOL)
