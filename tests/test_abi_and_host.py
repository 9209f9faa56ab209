"""CPU-side checks: the C-ABI library loads and exports every symbol that
include/skge_b200.h declares, the ctypes prototypes agree with the header, and
the host-side ranking logic (query flattening, filter lists, sharding, count
combination, regrouping) matches the oracle -- including a world_size-2 gloo run.
No compute call is made (there is no GPU here)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import cpu_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'skge_b200.h')

CTYPE = {'int': 'c_int', 'int64_t': 'c_long', 'float': 'c_float', 'size_t': 'c_ulong', 'uint64_t': 'c_ulong'}


def header_prototypes():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    protos = {}
    for m in re.finditer(r'SKGE_API\s+([\w\s\*]+?)\s*\b(skge_\w+)\s*\(([^;]*?)\)\s*;', src, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        kinds = []
        if args and args != 'void':
            for a in args.split(','):
                a = ' '.join(a.split())
                if '*' in a or a.startswith('skge_stream_t'):
                    kinds.append('ptr')
                else:
                    kinds.append(CTYPE[a.replace('const ', '').split(' ')[0]])
        protos[name] = (ret, kinds)
    return protos


def test_library_exports_every_declared_symbol():
    from skge import _ext
    lib = _ext.load_library()
    protos = header_prototypes()
    assert len(protos) >= 29
    for name in protos:
        assert hasattr(lib, name), name
    assert set(protos) == set(_ext.SIGNATURES), set(protos) ^ set(_ext.SIGNATURES)
    assert lib.skge_version() == 100


def test_ctypes_prototypes_match_header():
    import ctypes as C
    from skge import _ext
    protos = header_prototypes()
    for name, (ret, kinds) in protos.items():
        res, args = _ext.SIGNATURES[name]
        assert len(args) == len(kinds), name
        for i, (a, k) in enumerate(zip(args, kinds)):
            if k == 'ptr':
                assert a is C.c_void_p, (name, i)
            else:
                assert a.__name__ == k, (name, i, a.__name__, k)


def test_compute_without_gpu_fails_loudly():
    from skge import _ext
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        _ext.lib()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'scikit-kge_b200', 'skge')
    for f in os.listdir(pkg):
        if f.endswith('.py'):
            src = open(os.path.join(pkg, f)).read()
            assert not re.search(r'^\s*(from|import)\s+oracle', src, flags=re.M), f


def _toy(seed=0, N=40, M=3, ntrue=160, ntest=24):
    rng = np.random.default_rng(seed)
    true = np.unique(np.stack([rng.integers(N, size=ntrue), rng.integers(N, size=ntrue),
                               rng.integers(M, size=ntrue)], 1), axis=0)
    test = true[rng.choice(len(true), ntest, replace=False)]
    return N, M, true, test


def _oracle_counts(kind_name, E, R, test, true, lo, hi):
    """Per-shard counts computed with the oracle's scorers (numpy)."""
    from skge.ranking import flatten_queries
    kind, given, rel, target = flatten_queries(test)
    prepare, scores_o, scores_s = orc._eval_hooks(kind_name, E, R)
    idx, tt = orc.build_filter_index(test, true)
    cnt = np.zeros((2, len(kind)), dtype=np.int32)
    for j in range(len(kind)):
        prepare(int(rel[j]))
        sc = (scores_o if kind[j] == 0 else scores_s)(int(given[j]), int(rel[j])).flatten()
        t = sc[target[j]]
        cnt[0, j] = np.sum(sc[lo:hi] > t)
        known = tt[int(rel[j])]['os' if kind[j] == 0 else 'ss'][int(given[j])]
        known = {e for e in known if e != target[j] and lo <= e < hi}
        cnt[1, j] = sum(1 for e in known if sc[e] > t)
    return cnt


def test_filter_pairs_match_oracle_index():
    from skge.ranking import flatten_queries, build_filter_pairs
    N, M, true, test = _toy()
    kind, given, rel, target = flatten_queries(test)
    pq, pe = build_filter_pairs(true, kind, given, rel, target)
    got = {}
    for q, e in zip(pq.tolist(), pe.tolist()):
        got.setdefault(q, set()).add(e)
    idx, tt = orc.build_filter_index(test, true)
    for j in range(len(kind)):
        known = tt[int(rel[j])]['os' if kind[j] == 0 else 'ss'][int(given[j])]
        want = {e for e in known if e != target[j]}
        assert got.get(j, set()) == want
    assert (np.diff(pq.numpy()) >= 0).all()


@pytest.mark.parametrize('kind_name', ['transe', 'hole'])
@pytest.mark.parametrize('world', [1, 2, 3, 8])
def test_sharded_counts_recombine_to_reference_ranks(kind_name, world):
    from skge.ranking import shard_range, ranks_from_counts, regroup
    N, M, true, test = _toy(seed=3)
    rng = np.random.default_rng(1)
    E, R = rng.normal(size=(N, 8)), rng.normal(size=(M, 8))
    total = np.zeros((2, 2 * len(test)), dtype=np.int32)
    covered = 0
    for r in range(world):
        lo, hi = shard_range(N, r, world)
        covered += hi - lo
        total += _oracle_counts(kind_name, E, R, test, true, lo, hi)
    assert covered == N
    raw, filt = ranks_from_counts(torch.from_numpy(total))
    pos, fpos = regroup(test, raw.numpy(), filt.numpy())
    rpos, rfpos = orc.rank_positions(kind_name, E, R, test, true, tie='argsort')
    assert pos == rpos and fpos == rfpos
    assert list(pos.keys()) == list(rpos.keys())


def test_regroup_plan_is_reusable_and_handles_edge_cases():
    """The evaluator computes the grouping of the test triples once (regroup_plan) and reuses it
    for every pass: same dicts as the one-shot call, relations in order of first appearance,
    triples of a relation in test order; relation ids >= 65536 and an empty test set work."""
    from skge.ranking import regroup, regroup_plan
    rng = np.random.default_rng(0)
    for M in (7, 70000):
        n = 500
        test = np.stack([rng.integers(100, size=n), rng.integers(100, size=n), rng.integers(M, size=n)], 1)
        plan = regroup_plan(test)
        for _ in range(2):
            raw = rng.integers(1, 100, size=2 * n).astype(np.int32)
            filt = rng.integers(1, 100, size=2 * n).astype(np.int32)
            pos, fpos = regroup(test, raw, filt, plan)
            want, fwant = {}, {}
            for i, (s, o, p) in enumerate(test.tolist()):
                for dst, arr in ((want, raw), (fwant, filt)):
                    d = dst.setdefault(p, {'head': [], 'tail': []})
                    d['tail'].append(int(arr[i]))
                    d['head'].append(int(arr[n + i]))
            assert pos == want and fpos == fwant
            assert list(pos.keys()) == list(want.keys())
            assert regroup(test, raw, filt) == (pos, fpos)
    assert regroup_plan(np.zeros((0, 3), dtype=np.int64)) is None
    assert regroup(np.zeros((0, 3), dtype=np.int64), np.zeros(0), np.zeros(0)) == ({}, {})


def test_lazy_relation_ranks_behave_like_plain_dicts():
    """regroup hands out per-relation dicts whose lists are built on first use: every read path
    (indexing, iteration, equality in both directions, pickling, the reference's metric loop of
    skge/base.py:1086-1103) must see a plain ``{'head': [...], 'tail': [...]}``."""
    import pickle
    from skge.ranking import regroup, ranking_scores, _RelRanks
    test = np.array([(0, 1, 3), (2, 3, 1), (4, 5, 3)])
    raw = np.array([1, 2, 3, 4, 5, 6], dtype=np.int32)
    pos, fpos = regroup(test, raw, raw)
    want = {3: {'head': [4, 6], 'tail': [1, 3]}, 1: {'head': [5], 'tail': [2]}}
    assert pos == want and want == pos and not (pos != want)
    pos, _ = regroup(test, raw, raw)
    assert isinstance(pos[3], _RelRanks) and pos[3]['head'] == [4, 6] and type(pos[3]['head'][0]) is int
    pos, _ = regroup(test, raw, raw)
    assert sorted(pos[1].keys()) == ['head', 'tail'] and len(pos[1]) == 2 and 'tail' in pos[1]
    pos, _ = regroup(test, raw, raw)
    assert pickle.loads(pickle.dumps(pos)) == want and type(pickle.loads(pickle.dumps(pos))[3]) is dict
    pos, _ = regroup(test, raw, raw)
    assert dict(pos[3].items()) == want[3] and pos[3].get('nope') is None
    with pytest.raises(KeyError):
        pos[3]['nope']
    pos, fpos = regroup(test, raw, raw)
    assert ranking_scores(None, pos, fpos, 0, 'x') == pytest.approx(np.mean(1.0 / raw))
    pos, _ = regroup(test, raw, raw)
    pos[3]['head'].append(9)
    assert pos[3] == {'head': [4, 6, 9], 'tail': [1, 3]}


_GLOO_WORKER = r'''
import os, sys
sys.path[:0] = [%(root)r, os.path.join(%(root)r, 'scikit-kge_b200'), os.path.join(%(root)r, 'tests')]
import numpy as np, torch, torch.distributed as dist
from test_abi_and_host import _toy, _oracle_counts
from skge.ranking import shard_range, allreduce_counts, ranks_from_counts, regroup
from oracle import cpu_oracle as orc
dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%(port)d', rank=int(sys.argv[1]), world_size=2)
N, M, true, test = _toy(seed=5)
rng = np.random.default_rng(2)
E, R = rng.normal(size=(N, 8)), rng.normal(size=(M, 8))
lo, hi = shard_range(N, dist.get_rank(), 2)
cnt = torch.from_numpy(_oracle_counts('hole', E, R, test, true, lo, hi))
cnt = allreduce_counts(cnt)
raw, filt = ranks_from_counts(cnt)
pos, fpos = regroup(test, raw.numpy(), filt.numpy())
rpos, rfpos = orc.rank_positions('hole', E, R, test, true, tie='argsort')
assert pos == rpos and fpos == rfpos, 'rank %%d mismatch' %% dist.get_rank()
dist.barrier()
dist.destroy_process_group()
print('ok')
'''


def test_two_rank_gloo_count_reduction(tmp_path):
    port = 29500 + os.getpid() % 2000
    script = tmp_path / 'worker.py'
    script.write_text(_GLOO_WORKER % dict(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0 and 'ok' in o, o


def test_trainer_batch_bounds_and_api_surface():
    import skge
    from skge.base import StochasticTrainer
    assert set(['HolE', 'RESCAL', 'TransE', 'StochasticTrainer', 'PairwiseStochasticTrainer',
                'activation_functions']) <= set(dir(skge))
    assert set(skge.activation_functions) == {'linear', 'sigmoid', 'tanh', 'relu', 'softplus'}
    t = StochasticTrainer.__new__(StochasticTrainer)
    t.nbatches = 100
    b = t._batch_bounds(141442)        # WN18: 100 x 1414 + remainder 42 (skge/base.py:1264)
    assert len(b) == 101 and b[0] == (0, 1414) and b[-1] == (141400, 141442) and t.batch_size == 1414
    assert len(t._batch_bounds(1000)) == 100
    from skge.param import SGD, AdaGrad, normalize, normless1  # noqa: F401  (README.md:95)
    from skge.util import ccorr, cconv, grad_sum_matrix, unzip_triples
    ss, ps, os_ = unzip_triples([((1, 2, 3), 1.0), ((4, 5, 6), -1.0)])
    assert ss.tolist() == [1, 4] and ps.tolist() == [3, 6] and os_.tolist() == [2, 5]
    u, Sm, n = grad_sum_matrix([1, 2, 6, 4, 2, 3, 2])
    assert u.tolist() == [1, 2, 3, 4, 6] and n.flatten().tolist() == [1, 3, 1, 1, 1]
    a = np.arange(6.0).reshape(2, 3)
    np.testing.assert_allclose(ccorr(a, a[::-1]), orc.ccorr_direct(a, a[::-1]))
    np.testing.assert_allclose(cconv(a, a[::-1]), orc.cconv_direct(a, a[::-1]))


def test_table_checksum_detects_any_change_on_both_paths():
    """ranking._checksum validates everything cached across passes (packed shadows, norm bounds): the
    int64-word fast path and the int32 fallback (odd element count / unaligned view) must both change when
    a single element changes, and must not depend on a widening copy of the table."""
    import torch
    from skge import ranking
    g = torch.Generator().manual_seed(3)
    for shape, view in (((64, 6), slice(None)), ((7, 3), slice(None)), ((64, 6), slice(1, None))):
        t = torch.randn(*shape, generator=g)[view]
        a = ranking._checksum(t)
        assert a == ranking._checksum(t.clone())
        u = t.clone()
        u[-1, -1] = torch.nextafter(u[-1, -1], torch.tensor(10.0))
        assert ranking._checksum(u) != a
        u = t.clone()
        u[0, 0] = -u[0, 0]
        assert ranking._checksum(u) != a


def test_spectral_row_lengths_and_mixed_radix_plan():
    """Row lengths HolE's frequency-domain step takes (csrc/fft.cuh::spectral_len_ok, skge/hole.py), and the
    stage plan of the warp transform (a radix-2 stage if needed, radix-4, radix-3, radix-5 Stockham stages)
    restated in numpy against numpy's FFT -- the reference transforms any length (skge/util.py:27,50)."""
    from skge.hole import spectral_len_ok
    ok = [d for d in range(1, 1100) if spectral_len_ok(d)]
    assert {32, 64, 100, 128, 150, 200, 256, 300, 512, 1000, 1024} <= set(ok)
    assert not any(spectral_len_ok(d) for d in (10, 16, 24, 30, 33, 44, 70, 98, 149, 151, 154, 1026, 2048))
    for d in ok:
        h = d // 2
        for r in (2, 3, 5):
            while h % r == 0:
                h //= r
        assert d % 2 == 0 and 32 <= d <= 1024 and h == 1
    src = open(os.path.join(ROOT, 'scikit-kge_b200', 'csrc', 'fft.cuh')).read()
    assert 'd >= 32 && d <= 1024 && (d & 1) == 0 && fft_len_ok(d / 2)' in src      # the same rule on the device side

    def plan(n):
        m, n2, st = n, 0, []
        while m % 2 == 0:
            m //= 2
            n2 += 1
        st += [2] * (n2 & 1) + [4] * (n2 // 2)
        for r in (3, 5):
            while m % r == 0:
                st.append(r)
                m //= r
        assert m == 1
        return st

    def stockham(x, inverse):
        n, a, Ns = len(x), np.asarray(x, dtype=complex).copy(), 1
        sg = 1j if inverse else -1j
        for r in plan(n):
            out, q = np.zeros(n, complex), n // r
            for j in range(q):
                k = j % Ns
                j0 = (j - k) * r + k
                v = [a[j + t * q] * np.exp(sg * 2 * np.pi * t * k / (r * Ns)) for t in range(r)]
                for s_ in range(r):
                    out[j0 + s_ * Ns] = sum(v[t] * np.exp(sg * 2 * np.pi * s_ * t / r) for t in range(r))
            a, Ns = out, Ns * r
        return a

    rng = np.random.default_rng(5)
    for n in (16, 32, 75, 50, 125, 48, 27, 150):
        x = rng.normal(size=n) + 1j * rng.normal(size=n)
        np.testing.assert_allclose(stockham(x, False), np.fft.fft(x), atol=1e-10)
        np.testing.assert_allclose(stockham(x, True), np.fft.ifft(x) * n, atol=1e-10)
