#!/usr/bin/env python
"""Diagnostic: are the small configs host- or device-bound through fit()?
SKGE_INSTRUMENT=1 splits the host side of a graph-replayed minibatch (signature check, index copy, cudaGraphLaunch).
Measured on B200: config 1: 1.6 + 6.0 + 19.2 us of host work per 92 us minibatch, config 3: 1.4 + 5.6 + 57.5 us per
187 us -- device-bound.  (The plain mode's "host issue" time equals the total because every epoch ends with an
.item() read of the violation / loss counter before the callbacks run.)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'scikit-kge_b200')]
import numpy as np, torch
import skge
from skge.param import AdaGrad, SGD
from skge.sample import RandomModeSampler
from skge.synth import make_graph


def run(name, shape, model, d, pairwise, upd, margin=None, epochs=6):
    g = make_graph(shape, device='cuda')
    xs = g['train'].cpu().numpy()
    N, M = g['N'], g['M']
    m = {'transe': skge.TransE, 'hole': skge.HolE, 'rescal': skge.RESCAL}[model]((N, N, M), d)
    m.track_counters = False
    smp = RandomModeSampler(1, [0, 1], xs, (N, N, M))
    issue, total = [], []

    def cb(t):
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        issue.append(t1 - t.epoch_start)
        total.append(t2 - t.epoch_start)
        return True
    kw = dict(nbatches=100, max_epochs=epochs, learning_rate=0.1, samplef=smp.sample, param_update=upd, post_epoch=[cb])
    trn = skge.PairwiseStochasticTrainer(m, margin=margin, **kw) if pairwise else skge.StochasticTrainer(m, **kw)
    trn.fit(xs, np.ones(len(xs)))
    print('%s: host issue %.0f us / minibatch, until the device is done %.0f us / minibatch (median of epochs 2..)'
          % (name, np.median(issue[1:]) * 1e6 / 101, np.median(total[1:]) * 1e6 / 101), flush=True)


def instrument():
    """Split _GraphedStep.__call__ into its host-side pieces (signature check, index copy, graph launch)."""
    from skge import base, kernels
    acc = dict(sig=0.0, copy=0.0, replay=0.0, n=0)

    def call(self, batch):
        t0 = time.perf_counter()
        if self.signature is not None:
            sig = self.signature()
            if sig != self._sig:
                self._sig, self.slots = sig, {}
        slot = self.slots.get(batch.numel())
        if slot is None or slot['graph'] is None:
            return orig(self, batch)
        t1 = time.perf_counter()
        slot['idx'].copy_(batch)
        t2 = time.perf_counter()
        slot['graph'].replay()
        t3 = time.perf_counter()
        kernels.LAUNCHES['n'] += slot['launches']
        acc['sig'] += t1 - t0; acc['copy'] += t2 - t1; acc['replay'] += t3 - t2; acc['n'] += 1
    orig = base._GraphedStep.__call__
    base._GraphedStep.__call__ = call
    return acc


if __name__ == '__main__':
    if os.environ.get('SKGE_INSTRUMENT'):
        acc = instrument()
        run('cfg1', 'wn18', 'transe', 50, True, AdaGrad, 2.0)
        print('cfg1 per replayed minibatch: signature %.1f us, index copy %.1f us, graph launch %.1f us (host)'
              % tuple(acc[k] * 1e6 / acc['n'] for k in ('sig', 'copy', 'replay')))
        for k in ('sig', 'copy', 'replay', 'n'): acc[k] = 0
        run('cfg3', 'wn18', 'rescal', 100, False, SGD)
        print('cfg3 per replayed minibatch: signature %.1f us, index copy %.1f us, graph launch %.1f us (host)'
              % tuple(acc[k] * 1e6 / acc['n'] for k in ('sig', 'copy', 'replay')))
        sys.exit(0)
    run('cfg1', 'wn18', 'transe', 50, True, AdaGrad, 2.0)
    run('cfg2', 'wn18', 'hole', 150, True, AdaGrad, 0.2)
    run('cfg3', 'wn18', 'rescal', 100, False, SGD)
    run('cfg4', 'fb15k', 'transe', 200, True, AdaGrad, 2.0)
