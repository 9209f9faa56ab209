"""Debug: GPU ranks of sampled config-5 queries vs the unmodified reference (all engines)."""
import json, os, sys, subprocess, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'scikit-kge_b200')]
import numpy as np, torch
import bench
class A: workload='cfg5'; test_triples=0
w = bench.build_workload(A)
import skge
from skge import ranking
dev = torch.device('cuda', 0)
N, M, d = w['N'], w['M'], w['d']
mdl = skge.HolE((N, N, M), d)
mdl.E.data.copy_(w['E'].to(dev)); mdl.R.data.copy_(w['R'].to(dev))
test, true = w['test'].numpy(), w['true'].numpy()
E, R = w['E'].numpy(), w['R'].numpy()
samples = [bench.pick_sample(test, true, 3 + i) for i in range(2)]
with tempfile.TemporaryDirectory() as tmp:
    npz, out = os.path.join(tmp, 'w.npz'), os.path.join(tmp, 'ranks.json')
    arrs = dict(N=N, M=M, d=d, model='hole', desc=w['desc'], te=w['te'], E=E, R=R, nsamples=2, nrel_test=1000)
    for i, (t, tr) in enumerate(samples):
        arrs['test%d' % i], arrs['true%d' % i] = t, tr
    np.savez(npz, **arrs)
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--from-npz', npz, '--emit-ranks', out], capture_output=True, text=True)
    print(r.stderr[-500:])
    ref = json.load(open(out))['ranks']
print('ref', ref)
for eng, ns in (('sweep', 0), ('umma', 3), ('umma', 2), ('umma', 1)):
    got = []
    for t, tr in samples:
        ev = ranking.HolEEval(t, tr); ev.engine = eng
        if ns: ev.nsplit = ns
        pos, fpos = ev.positions(mdl)
        got.append([(pos[p]['tail'][i], fpos[p]['tail'][i], pos[p]['head'][i], fpos[p]['head'][i]) for p in pos for i in range(len(pos[p]['tail']))])
    print(eng, ns, got, ev.last_stats)
for t, tr in samples:
    print('test', t.tolist(), 'ntrue', len(tr))
