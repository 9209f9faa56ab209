"""The CPU baseline of bench.py is the UNMODIFIED reference shipped as oracle/_ref (oracle/build_ref.py).
These tests check the copy against its manifest (and against /root/reference where that exists), and that
oracle/ref_loader.py runs the reference's own evaluator with results equal to the oracle's."""
import hashlib
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, 'oracle', '_ref')

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'skge')),
                                reason='oracle/_ref not built (python oracle/build_ref.py where /root/reference exists)')


def test_reference_copy_matches_its_manifest_and_the_source():
    man = json.load(open(os.path.join(REF, 'MANIFEST.json')))
    assert len(man['files']) >= 10
    for rel, digest in man['files'].items():
        data = open(os.path.join(REF, rel), 'rb').read()
        assert hashlib.sha256(data).hexdigest() == digest, rel
        src = os.path.join('/root/reference', rel)
        if os.path.exists(src):                         # build container only
            assert open(src, 'rb').read() == data, rel
    # nothing but the reference's own files and the manifest
    extra = set(os.listdir(os.path.join(REF, 'skge'))) - {os.path.basename(k) for k in man['files']} - {'__pycache__'}
    assert not extra, extra


def test_reference_evaluator_runs_under_the_loader_and_agrees_with_the_oracle():
    code = r'''
import sys, json
sys.path.insert(0, %r)
import numpy as np
from oracle import ref_loader
ref = ref_loader.load()
from oracle import cpu_oracle as orc
rng = np.random.default_rng(0)
N, M, d = 60, 3, 16
E = rng.normal(size=(N, d)); R = rng.normal(size=(M, d))
true = np.unique(np.stack([rng.integers(N, size=300), rng.integers(N, size=300), rng.integers(M, size=300)], 1), axis=0)
test = true[:25]
m = ref.HolE((N, N, M), d)
m.E[...] = E; m.R[...] = R
ev = ref.HolEEval([tuple(map(int, t)) for t in test], [tuple(map(int, t)) for t in true])
pos, fpos = ev.positions(m)
opos, ofpos = orc.rank_positions('hole', E, R, test, true, tie='argsort')
ok = all(list(pos[p][s]) == list(opos[p][s]) and list(fpos[p][s]) == list(ofpos[p][s]) for p in opos for s in ('head', 'tail'))
print(json.dumps({'ok': bool(ok), 'n': int(sum(len(v['head']) for v in pos.values())), 'root': ref.root}))
''' % ROOT
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-800:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out['ok'] and out['n'] == 25
    assert os.path.realpath(out['root']) == os.path.realpath(REF)
