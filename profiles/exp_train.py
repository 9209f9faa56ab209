#!/usr/bin/env python
"""Diagnostic: config-5-sized fused minibatch steps (P = 1M pairs, N = 1M, d = 256 or argv[3]), timed with CUDA events.
SKGE_SPECTRAL=0 forces HolE's per-pair kernels (direct O(d^2) correlations for d that is not a power of two).
Run under `ncu --metrics gpu__time_duration.sum` for the per-kernel breakdown."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'scikit-kge_b200')]
import torch
import skge
from skge.param import AdaGrad
from skge.sample import RandomModeSampler
from skge.synth import make_graph

def main(model='hole', N=1000000, M=1000, d=256, B=500000, steps=3):
    dev = torch.device('cuda')
    g = make_graph((N, M, B * 4, 1, 1), device=dev, seed=3)
    xs = g['train'].cpu().numpy()
    m = (skge.HolE if model == 'hole' else skge.TransE)((N, N, M), d)
    m.track_counters = False
    smp = RandomModeSampler(1, [0, 1], xs, (N, N, M))
    trn = skge.PairwiseStochasticTrainer(m, nbatches=4, margin=0.2 if model == 'hole' else 2.0, max_epochs=1,
                                         learning_rate=0.1, samplef=smp.sample, param_update=AdaGrad)
    trn._setup_fused()
    smp.ensure_device()
    if hasattr(m, '_prepare_fused') and os.environ.get('SKGE_SPECTRAL', '1') == '1':
        m._prepare_fused()      # HolE: frequency-domain training state
    perm = torch.randperm(len(xs), device=dev).to(torch.int32)
    for it in range(steps + 1):
        batch = perm[(it % 4) * B:(it % 4 + 1) * B]
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record()
        pos, neg, valid = smp.device_sample(batch, B, it)
        b.record()
        m._fused_pair_step(trn._updaters, pos, neg, valid, trn._counts, trn._nviol_dev)
        c.record()
        torch.cuda.synchronize()
        nv, ue, ur, _ = trn._counts.tolist()
        P = 2 * B
        byts = 4 * d * (4 * P + 4 * (ue + ur)) + 24 * P
        print('%s step %d: sample %.2f ms, step %.2f ms | nviol %d U_E %d U_R %d | algorithmic %.2f GB -> %.0f GB/s, %.2f M triples/s'
              % (model, it, a.elapsed_time(b), b.elapsed_time(c), nv, ue, ur, byts / 1e9,
                 byts / 1e6 / a.elapsed_time(c), B / 1e3 / a.elapsed_time(c)), flush=True)

if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else 'hole', steps=int(sys.argv[2]) if len(sys.argv) > 2 else 3,
         d=int(sys.argv[3]) if len(sys.argv) > 3 else 256)
