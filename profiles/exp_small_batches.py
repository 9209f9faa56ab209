#!/usr/bin/env python
"""Diagnostic: where does the time of a small minibatch go (configs 1-4: <= 10k pairs per batch)?
Host time per fused batch (no sync) vs device time per batch."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'scikit-kge_b200')]
import torch
import skge
from skge.param import AdaGrad
from skge.sample import RandomModeSampler
from skge.synth import make_graph

def main(model, shape, d):
    dev = torch.device('cuda')
    g = make_graph(shape, device=dev)
    xs = g['train'].cpu().numpy()
    N, M = g['N'], g['M']
    m = (skge.HolE if model == 'hole' else skge.TransE)((N, N, M), d)
    m.track_counters = False
    smp = RandomModeSampler(1, [0, 1], xs, (N, N, M))
    trn = skge.PairwiseStochasticTrainer(m, nbatches=100, margin=0.2 if model == 'hole' else 2.0, max_epochs=1,
                                         learning_rate=0.1, samplef=smp.sample, param_update=AdaGrad)
    trn._setup_fused()
    smp.ensure_device()
    n = len(xs)
    bounds = trn._batch_bounds(n)
    for rep in range(3):
        perm = torch.randperm(n, device=dev).to(torch.int32)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i, (lo, hi) in enumerate(bounds):
            pos, neg, valid = smp.device_sample(perm[lo:hi], hi - lo, rep * 1000 + i)
            m._fused_pair_step(trn._updaters, pos, neg, valid, trn._counts, trn._nviol_dev)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print('%s %s d=%d: %d batches: host issue %.1f ms (%.0f us/batch), total %.1f ms (%.0f us/batch) -> %.2f M triples/s'
              % (model, shape, d, len(bounds), (t1 - t0) * 1e3, (t1 - t0) * 1e6 / len(bounds), (t2 - t0) * 1e3,
                 (t2 - t0) * 1e6 / len(bounds), n / (t2 - t0) / 1e6), flush=True)

if __name__ == '__main__':
    main('transe', 'wn18', 50)
    main('hole', 'wn18', 150)
    main('transe', 'fb15k', 200)
