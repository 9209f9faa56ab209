"""GPU parity of the filtered-ranking pass: golden vectors of the reference,
the oracle on seeded graphs (incl. RESCAL, which has no reference evaluator),
query chunking, emulated entity sharding and the per-query hooks."""
import numpy as np
import pytest
import torch

from oracle import cpu_oracle as orc

pytestmark = pytest.mark.gpu

# CTA groupings of the refine kernel under test (SKGE_TEST_CG=1 isolates the single-CTA kernel)
import os
CGS = [int(x) for x in os.environ.get('SKGE_TEST_CG', '1,2').split(',')]
# large-sweep engines: 'refine' = two tensor-core products + one refined cross term (csrc/rank_refine.cu),
# 'single' = one product + both cross terms refined (csrc/rank_single.cu)
MODES = os.environ.get('SKGE_TEST_MODES', 'refine,single').split(',')


def _set_engine(ev, mode, cg):
    """Selects a large-sweep engine on the evaluator; returns the engine name last_stats must report."""
    ev.cta_group = cg
    if mode == 'single':
        ev.engine, ev.nsplit = 'single', 0
        return 'tcgen05-f16x1-refined'
    ev.engine, ev.nsplit = 'umma', 2
    return 'tcgen05-f16x2'


def _flat(d, rel, side):
    return np.concatenate([np.array(d[int(p)][side]) for p in rel])


def _model(kind, E0, R0):
    import skge
    N, d = E0.shape
    M = R0.shape[0]
    cls = {'transe': skge.TransE, 'hole': skge.HolE, 'rescal': skge.RESCAL}[kind]
    m = cls((N, N, M), d)
    m.E[...] = E0
    (m.W if kind == 'rescal' else m.R)[...] = R0
    return m


def _evaluator(kind):
    from skge.ranking import TransEEval, HolEEval, RESCALEval
    return {'transe': TransEEval, 'hole': HolEEval, 'rescal': RESCALEval}[kind]


@pytest.mark.parametrize('name,kind', [('rank_transe', 'transe'), ('rank_hole', 'hole')])
def test_positions_match_reference_golden(golden, name, kind):
    from skge.ranking import ranking_scores, compute_scores
    g = golden(name)
    m = _model(kind, g['E0'], g['R0'])
    ev = _evaluator(kind)([tuple(t) for t in g['test'].tolist()], [tuple(t) for t in g['true'].tolist()])
    pos, fpos = ev.positions(m)
    assert list(pos.keys()) == [int(p) for p in g['rel']]
    for side in ('head', 'tail'):
        np.testing.assert_array_equal(_flat(pos, g['rel'], side), g['pos_' + side])
        np.testing.assert_array_equal(_flat(fpos, g['rel'], side), g['fpos_' + side])
    allpos = np.concatenate([_flat(pos, g['rel'], 'head'), _flat(pos, g['rel'], 'tail')])
    allf = np.concatenate([_flat(fpos, g['rel'], 'head'), _flat(fpos, g['rel'], 'tail')])
    np.testing.assert_allclose(compute_scores(allpos), g['raw'], rtol=1e-12)
    np.testing.assert_allclose(compute_scores(allf), g['filt'], rtol=1e-12)
    for k in (1, 3, 10):
        assert compute_scores(allf, hits=k)[2] == pytest.approx(orc.compute_scores(g['fpos_head'].tolist()
                                                                + g['fpos_tail'].tolist(), hits=k)[2])
    assert ranking_scores(None, pos, fpos, 1, 'TEST') == pytest.approx(float(g['filt'][0]))


def test_appendix_a4_untied_ranks(golden):
    g = golden('appendix_a')
    for tag, kind in (('te', 'transe'), ('ho', 'hole')):
        m = _model(kind, g['E0'], g['R0'])
        ev = _evaluator(kind)(g['a4_test'], g['a4_true'])
        pos, fpos = ev.positions(m)
        _, _, margins = orc.rank_positions(kind, g['E0'], g['R0'], g['a4_test'], g['a4_true'], with_scores=True)
        for p in pos:
            for side in ('head', 'tail'):
                for i, mg in enumerate(margins[p][side]):
                    if mg > 1e-6:      # ties with the target are tie-order dependent in the reference
                        assert pos[p][side][i] == int(g['a4_%s_pos_%d_%s' % (tag, p, side)][i])
                        assert fpos[p][side][i] == int(g['a4_%s_fpos_%d_%s' % (tag, p, side)][i])


@pytest.mark.parametrize('mode', MODES)
@pytest.mark.parametrize('cg', CGS)
def test_refine_engine_matches_reference_golden(golden, cg, mode):
    """The engine bench.py times (two tensor-core products + int8 refinement, both CTA groupings)
    on the reference's own ranks (skge/base.py:913-1031 run through oracle/make_golden.py)."""
    g = golden('rank_hole')
    m = _model('hole', g['E0'], g['R0'])
    ev = _evaluator('hole')([tuple(t) for t in g['test'].tolist()], [tuple(t) for t in g['true'].tolist()])
    name = _set_engine(ev, mode, cg)
    pos, fpos = ev.positions(m)
    assert ev.last_stats['engine'] == name
    assert list(pos.keys()) == [int(p) for p in g['rel']]
    for side in ('head', 'tail'):
        np.testing.assert_array_equal(_flat(pos, g['rel'], side), g['pos_' + side])
        np.testing.assert_array_equal(_flat(fpos, g['rel'], side), g['fpos_' + side])


@pytest.mark.parametrize('cg', CGS)
@pytest.mark.parametrize('kind,N,d', [('hole', 1500, 150), ('hole', 900, 256), ('rescal', 800, 100), ('hole', 333, 37),
                                      ('hole', 129, 64), ('hole', 2100, 200)])
@pytest.mark.parametrize('mode', MODES)
def test_refine_engine_matches_oracle_on_random_graphs(kind, N, d, cg, mode):
    """Same cases as the three-product engine: ragged query chunks, ragged entity tiles, odd and
    even numbers of 64-wide k chunks (the int8 query rows are swizzled only for even counts)."""
    M = 5
    rng, true, test = _graph(N + d, N, M, 6 * N, 150)
    E0 = (rng.normal(size=(N, d)) * 0.3).astype(np.float32).astype(np.float64)
    shape = (M, d, d) if kind == 'rescal' else (M, d)
    R0 = (rng.normal(size=shape) * 0.3).astype(np.float32).astype(np.float64)
    m = _model(kind, E0, R0)
    ev = _evaluator(kind)(test, true)
    ev.chunk_queries = 128
    name = _set_engine(ev, mode, cg)
    pos, fpos = ev.positions(m)
    assert ev.last_stats['engine'] == name
    opos, ofpos, margins = orc.rank_positions(kind, E0, R0, test, true, tie='argsort', with_scores=True)
    assert list(pos.keys()) == list(opos.keys())
    nt = 0
    for p in opos:
        for side in ('head', 'tail'):
            for i, mg in enumerate(margins[p][side]):
                if mg > 1e-6:
                    nt += 1
                    assert pos[p][side][i] == opos[p][side][i], (p, side, i)
                    assert fpos[p][side][i] == ofpos[p][side][i], (p, side, i)
    assert nt >= 290


@pytest.mark.parametrize('mode', MODES)
@pytest.mark.parametrize('cg', CGS)
def test_refine_engine_on_a_million_entities(cg, mode):
    """A 1 M-entity x 2 k-query slice of the benchmarked workload (HolE d = 256): the refine
    engine's counts equal the fp32 sweep engine's bit for bit, and so do 2 emulated shards."""
    N, M, d, te = 1000000, 1000, 256, 1000
    gen = torch.Generator(device='cuda').manual_seed(5)
    import skge
    m = skge.HolE((N, N, M), d)
    m.E.data.copy_(torch.randn(N, d, device='cuda', generator=gen) / d ** 0.5)
    m.R.data.copy_(torch.randn(M, d, device='cuda', generator=gen) / d ** 0.5)
    rng = np.random.default_rng(9)
    test = np.stack([rng.integers(N, size=te), rng.integers(N, size=te), rng.integers(M, size=te)], 1)
    true = np.concatenate([test, np.stack([test[:, 0], rng.integers(N, size=te), test[:, 2]], 1)])
    ev = _evaluator('hole')(test, true)
    ev.engine = 'sweep'
    ref = ev.count_pass(m)
    name = _set_engine(ev, mode, cg)
    got = ev.count_pass(m)
    assert ev.last_stats['engine'] == name
    assert torch.equal(got, ref)
    parts = sum(ev.count_pass(m, world=(r, 2)) for r in range(2))
    assert torch.equal(parts, ref)


def _graph(seed, N, M, ntrue, ntest):
    rng = np.random.default_rng(seed)
    true = np.unique(np.stack([rng.integers(N, size=ntrue), rng.integers(N, size=ntrue),
                               rng.integers(M, size=ntrue)], 1), axis=0)
    test = true[rng.choice(len(true), ntest, replace=False)]
    return rng, true, test


@pytest.mark.parametrize('kind,N,d,engine', [
    ('transe', 1500, 50, 'sweep'), ('transe', 700, 200, 'sweep'), ('transe', 257, 7, 'sweep'),
    ('hole', 1500, 150, 'sweep'), ('hole', 900, 256, 'sweep'), ('rescal', 800, 100, 'sweep'), ('hole', 333, 37, 'sweep'),
    ('hole', 1500, 150, 'umma'), ('hole', 900, 256, 'umma'), ('rescal', 800, 100, 'umma'), ('hole', 333, 37, 'umma'),
    ('hole', 129, 64, 'umma'), ('hole', 700, 300, 'auto')])
def test_positions_match_oracle_on_random_graphs(kind, N, d, engine):
    M = 5
    rng, true, test = _graph(N + d, N, M, 6 * N, 150)
    E0 = (rng.normal(size=(N, d)) * 0.3).astype(np.float32).astype(np.float64)
    shape = (M, d, d) if kind == 'rescal' else (M, d)
    R0 = (rng.normal(size=shape) * 0.3).astype(np.float32).astype(np.float64)
    m = _model(kind, E0, R0)
    ev = _evaluator(kind)(test, true)
    ev.chunk_queries = 128        # several chunks, the last one ragged
    ev.engine = engine
    pos, fpos = ev.positions(m)
    want = {'sweep': 'fp32-sweep', 'umma': 'tcgen05-f16x3', 'auto': 'fp32-sweep' if d > 256 else 'tcgen05-f16x3'}
    assert ev.last_stats['engine'] == want[engine]
    opos, ofpos, margins = orc.rank_positions(kind, E0, R0, test, true, tie='argsort', with_scores=True)
    assert list(pos.keys()) == list(opos.keys())
    nt = 0
    for p in opos:
        for side in ('head', 'tail'):
            for i, mg in enumerate(margins[p][side]):
                if mg > 1e-6:
                    nt += 1
                    assert pos[p][side][i] == opos[p][side][i], (p, side, i)
                    assert fpos[p][side][i] == ofpos[p][side][i], (p, side, i)
    assert nt >= 290
    assert ev.last_stats['filter_pairs'] > 0


@pytest.mark.parametrize('cg', CGS)
@pytest.mark.parametrize('N,d,te', [(40943, 150, 700), (20000, 256, 1500), (5000, 64, 300)])
def test_tensor_core_engine_counts_equal_the_fp32_engine(N, d, te, cg):
    """Both coarse engines settle their undecided band in fp64, so the final counts
    must agree exactly (config-2-sized table and a d = 256 table); the single-product
    fp16 mode (nsplit = 1) must agree too, it only lists more candidates."""
    M = 18
    rng, true, test = _graph(N, N, M, 3 * N, te)
    E0 = (rng.normal(size=(N, d)) / np.sqrt(d)).astype(np.float32)
    R0 = (rng.normal(size=(M, d)) / np.sqrt(d)).astype(np.float32)
    m = _model('hole', E0, R0)
    ev = _evaluator('hole')(test, true)
    ev.engine = 'sweep'
    ref = ev.count_pass(m)
    ev.engine = 'umma'
    ev.nsplit = 3
    got = ev.count_pass(m)
    c3 = ev.last_stats['candidates']
    assert torch.equal(got, ref)
    ev.nsplit = 1
    got1 = ev.count_pass(m)
    assert torch.equal(got1, ref)
    assert ev.last_stats['candidates'] > c3 and ev.last_stats['engine'] == 'tcgen05-f16x1'
    # two products on the tensor cores, the third added in the epilogue for the wide-band pairs:
    # the refined scores are tested against the same tight band, so the fp64 workload stays small
    ev.nsplit = 2
    ev.cta_group = cg
    got2 = ev.count_pass(m)
    assert torch.equal(got2, ref)
    assert ev.last_stats['engine'] == 'tcgen05-f16x2'
    assert ev.last_stats['candidates'] < 4 * max(c3, 64)
    # the shadow of the shard is cached behind the table's checksum: a second pass reuses it, a
    # changed table rebuilds it
    key = ev._engines[('umma', 2, cg)]._shadow_key
    assert torch.equal(ev.count_pass(m), ref) and ev._engines[('umma', 2, cg)]._shadow_key == key
    m.E.data[17] *= 1.5
    ev.engine = 'sweep'
    ref2 = ev.count_pass(m)
    ev.engine = 'umma'
    assert torch.equal(ev.count_pass(m), ref2) and ev._engines[('umma', 2, cg)]._shadow_key != key


@pytest.mark.parametrize('cg', CGS)
@pytest.mark.parametrize('mode', MODES)
@pytest.mark.parametrize('N,d', [(7777, 256), (3001, 96)])
def test_refine_mode_with_mixed_row_norms(N, d, cg, mode):
    """nsplit = 2 packs the shard by decreasing row norm and widens the band per 128-row tile.
    Rows spanning a factor 50 in norm, a ragged last tile and emulated shards must still give the
    fp32 engine's counts bit for bit, with candidates reported under their original ids."""
    M = 11
    rng, true, test = _graph(N, N, M, 3 * N, 400)
    E0 = (rng.normal(size=(N, d)) / np.sqrt(d) * 10.0 ** rng.uniform(-1.7, 0.0, size=(N, 1))).astype(np.float32)
    R0 = (rng.normal(size=(M, d)) / np.sqrt(d)).astype(np.float32)
    m = _model('hole', E0, R0)
    ev = _evaluator('hole')(test, true)
    ev.engine = 'sweep'
    ref = ev.count_pass(m)
    name = _set_engine(ev, mode, cg)
    got = ev.count_pass(m)
    assert ev.last_stats['engine'] == name
    assert torch.equal(got, ref)
    parts = sum(ev.count_pass(m, world=(r, 3)) for r in range(3))
    assert torch.equal(parts, ref)


@pytest.mark.parametrize('kind', ['transe', 'hole'])
@pytest.mark.parametrize('world', [2, 3, 8])
def test_emulated_entity_shards_sum_to_the_single_gpu_counts(kind, world):
    """Counts are integers: G shards must reproduce G = 1 bit for bit."""
    N, M, d = 1001, 4, 64
    rng, true, test = _graph(7, N, M, 5000, 120)
    E0 = (rng.normal(size=(N, d)) * 0.3).astype(np.float32)
    R0 = (rng.normal(size=(M, d)) * 0.3).astype(np.float32)
    m = _model(kind, E0, R0)
    ev = _evaluator(kind)(test, true)
    full = ev.count_pass(m, world=(0, 1))
    acc = torch.zeros_like(full)
    for r in range(world):
        acc += ev.count_pass(m, world=(r, world))
    assert torch.equal(acc, full)


def test_score_hooks_match_reference_scorers(golden):
    """scores_o / scores_s (skge/run_transe.py:20-29, skge/run_hole.py:15-19)."""
    for name, kind in (('rank_transe', 'transe'), ('rank_hole', 'hole')):
        g = golden(name)
        m = _model(kind, g['E0'], g['R0'])
        ev = _evaluator(kind)(g['test'], g['true'])
        prepare, so, ss = orc._eval_hooks(kind, g['E0'], g['R0'])
        for s, o, p in g['test'][:5].tolist():
            prepare(p)
            ev.prepare(m, p)
            np.testing.assert_allclose(ev.scores_o(m, s, p), so(s, p), rtol=1e-12, atol=1e-13)
            np.testing.assert_allclose(ev.scores_s(m, o, p), ss(o, p), rtol=1e-12, atol=1e-13)


def test_ranking_edge_cases():
    """Empty test set, queries without filter entries, duplicated true triples, a table whose
    size is not a multiple of any tile, a single entity shard smaller than a tile."""
    from skge.ranking import HolEEval, TransEEval
    rng = np.random.default_rng(3)
    N, M, d = 131, 2, 64
    E0 = (rng.normal(size=(N, d)) * 0.3).astype(np.float32).astype(np.float64)
    R0 = (rng.normal(size=(M, d)) * 0.3).astype(np.float32).astype(np.float64)
    for kind, Ev in (('hole', HolEEval), ('transe', TransEEval)):
        m = _model(kind, E0, R0)
        assert Ev(np.zeros((0, 3), dtype=np.int64), [(0, 1, 0)]).positions(m) == ({}, {})
        test = np.array([(5, 6, 0), (7, 8, 1), (5, 9, 0)])
        true = np.array([(5, 6, 0), (5, 6, 0), (7, 8, 1), (5, 9, 0), (5, 9, 0), (100, 6, 0)])   # duplicates
        got = Ev(test, true).positions(m)
        assert got == orc.rank_positions(kind, E0, R0, test, true, tie='count')
        ev = Ev(test, true)
        acc = sum(ev.count_pass(m, world=(r, 8)) for r in range(8))      # 17-row shards
        assert torch.equal(acc, ev.count_pass(m, world=(0, 1)))
