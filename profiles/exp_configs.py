#!/usr/bin/env python
"""Training throughput of BASELINE.json's configs 1-4 through the public trainer API
(steady-state epochs, on-device sampler), plus ranking of the RESCAL config."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'scikit-kge_b200')]
import numpy as np, torch
import skge
from skge.param import AdaGrad, SGD
from skge.sample import RandomModeSampler
from skge.synth import make_graph
from skge.ranking import RESCALEval, TransEEval, HolEEval

def run(name, shape, model, d, pairwise, upd, margin=None, epochs=6):
    g = make_graph(shape, device='cuda')
    xs = g['train'].cpu().numpy()
    N, M = g['N'], g['M']
    cls = {'transe': skge.TransE, 'hole': skge.HolE, 'rescal': skge.RESCAL}[model]
    m = cls((N, N, M), d)
    m.track_counters = False
    smp = RandomModeSampler(1, [0, 1], xs, (N, N, M))
    times = []
    def cb(t):
        torch.cuda.synchronize(); times.append(time.perf_counter() - t.epoch_start); return True
    kw = dict(nbatches=100, max_epochs=epochs, learning_rate=0.1, samplef=smp.sample, param_update=upd, post_epoch=[cb])
    trn = skge.PairwiseStochasticTrainer(m, margin=margin, **kw) if pairwise else skge.StochasticTrainer(m, **kw)
    trn.fit(xs, np.ones(len(xs)))
    med = float(np.median(times[1:]))
    Ev = {'transe': TransEEval, 'hole': HolEEval, 'rescal': RESCALEval}[model]
    ev = Ev(g['test'], torch.cat([g['train'], g['valid'], g['test']]))
    ev.positions(m)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): ev.positions(m)
    torch.cuda.synchronize(); rk = (time.perf_counter() - t0) / 3
    print('%s: %s d=%d  train %.1f ms/epoch (%.0f us/batch) = %.2f M triples/s | rank %d queries in %.1f ms = %.2f M queries/s (%s)'
          % (name, model, d, med * 1e3, med * 1e6 / 101, len(xs) / med / 1e6, 2 * len(g['test']), rk * 1e3,
             2 * len(g['test']) / rk / 1e6, ev.last_stats['engine']), flush=True)

if __name__ == '__main__':
    only = sys.argv[1:]          # e.g. `exp_configs.py cfg3` (under ncu: SKGE_EPOCHS=2)
    ep = int(os.environ.get('SKGE_EPOCHS', '6'))
    for cfg in (('cfg1', 'wn18', 'transe', 50, True, AdaGrad, 2.0), ('cfg2', 'wn18', 'hole', 150, True, AdaGrad, 0.2),
                ('cfg3', 'wn18', 'rescal', 100, False, SGD, None), ('cfg4', 'fb15k', 'transe', 200, True, AdaGrad, 2.0)):
        if not only or cfg[0] in only:
            run(*cfg, epochs=ep)
