"""Row f3: models pickled here load in the UNMODIFIED reference and vice versa.
Runs only where /root/reference exists (the build container); CPU only -- parameters are
merely held in host memory, no kernel runs."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'skge')), reason='reference not present')

SHIM = textwrap.dedent('''
    import sys, types, collections, collections.abc, logging, pickle
    sys.path[:0] = [%r, %r]
    sys.modules['trident'] = types.ModuleType('trident')
    collections.Hashable = collections.abc.Hashable
    import numpy as np; np.Inf = np.inf
    import warnings; warnings.simplefilter('ignore')
    import skge, skge.base as base
    logging.disable(logging.CRITICAL)
''') % (REF, os.path.join(REF, 'skge'))


def run_ref(code):
    r = subprocess.run([sys.executable, '-c', SHIM + textwrap.dedent(code)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_reference_loads_our_pickles(tmp_path):
    import skge
    from skge import activation_functions as afs
    from oracle import cpu_oracle as orc
    rng = np.random.default_rng(0)
    for name, m in (('transe', skge.TransE((9, 9, 3), 6, l1=False)),
                    ('hole', skge.HolE((9, 9, 3), 6, rparam=0.25, af=afs['tanh'])),
                    ('rescal', skge.RESCAL((9, 9, 3), 6))):
        for p in m.params.values():
            p[...] = rng.normal(size=p.shape).astype(np.float32)
        f = str(tmp_path / (name + '.pkl'))
        m.save(f)
        out = run_ref('''
            m = base.Model.load(%r)
            print(type(m).__module__, type(m).__name__, type(m.E).__module__, type(m.E).__name__, m.E.dtype, m.ncomp)
            ss, ps, os_ = np.array([0, 1, 2]), np.array([0, 1, 2]), np.array([3, 4, 5])
            print(' '.join('%%.17g' %% v for v in m._scores(ss, ps, os_)))
            print(sorted(m.hyperparams))
        ''' % f)
        lines = out.strip().splitlines()
        assert lines[0] == 'skge.%s %s skge.param Parameter float64 6' % (name, type(m).__name__)
        got = np.array([float(v) for v in lines[1].split()])
        E = np.asarray(m.E, dtype=np.float64)
        P2 = np.asarray(m.W if name == 'rescal' else m.R, dtype=np.float64)
        s, p, o = np.array([0, 1, 2]), np.array([0, 1, 2]), np.array([3, 4, 5])
        want = {'transe': lambda: orc.transe_scores(E, P2, s, p, o, False),
                'hole': lambda: orc.hole_scores(E, P2, s, p, o),
                'rescal': lambda: orc.rescal_scores(E, P2, s, p, o)}[name]()
        np.testing.assert_allclose(got, want, rtol=1e-12)
        assert lines[2] == str(sorted(m.hyperparams))


def test_we_load_reference_pickles(tmp_path):
    import skge
    from skge.base import Model
    from skge.param import normalize, normless1
    f = str(tmp_path / 'ref_hole.pkl')
    g = str(tmp_path / 'ref_transe.pkl')
    out = run_ref('''
        from skge import HolE, TransE
        np.random.seed(3)
        m = HolE((8, 8, 2), 5, rparam=0.5); m.save(%r)
        t = TransE((8, 8, 2), 5, l1=True); t.save(%r)
        print(' '.join('%%.17g' %% v for v in np.asarray(m.E).ravel()))
        print(' '.join('%%.17g' %% v for v in np.asarray(t.R).ravel()))
    ''' % (f, g))
    lines = out.strip().splitlines()
    m = Model.load(f)
    assert isinstance(m, skge.HolE) and m.rparam == 0.5 and m.ncomp == 5 and m.af is skge.actfun.Sigmoid
    np.testing.assert_allclose(np.asarray(m.E, dtype=np.float64).ravel(), [float(v) for v in lines[0].split()],
                               rtol=1e-6)
    assert m.E.post is normless1 and m.R.post is None and m.E.name == 'E'
    t = Model.load(g)
    assert isinstance(t, skge.TransE) and t.l1 is True and t.E.post is normalize
    np.testing.assert_allclose(np.asarray(t.R, dtype=np.float64).ravel(), [float(v) for v in lines[1].split()],
                               rtol=1e-6)
