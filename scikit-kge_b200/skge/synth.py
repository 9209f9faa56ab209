"""Synthetic knowledge graphs of the benchmark shapes.

There is no network, hence no WN18 / FB15k: the benchmark configurations are
reproduced as seeded synthetic graphs with the named entity / relation / triple
counts.  Entities are drawn Zipf-like (rank^-0.75), relations Zipf (rank^-1);
triples are de-duplicated and split train / valid / test.  Pure torch, so it
runs on the GPU (the 50M-triple graph takes a few seconds there) or on the CPU.
"""
import numpy as np
import torch

SEED = 20261018

#        N          M      train      valid   test
SHAPES = {
    'wn18':  (40943, 18, 141442, 5000, 5000),
    'fb15k': (14951, 1345, 483142, 50000, 59071),
    'syn1m': (1000000, 1000, 50000000, 100000, 100000),
}


def _zipf_sampler(n, alpha, device):
    w = torch.arange(1, n + 1, device=device, dtype=torch.float64) ** (-alpha)
    cdf = torch.cumsum(w, 0)
    cdf /= cdf[-1].clone()
    return cdf


def _draw(cdf, k, gen, perm):
    u = torch.rand(k, device=cdf.device, dtype=torch.float64, generator=gen)
    r = torch.searchsorted(cdf, u).clamp_(max=cdf.numel() - 1)
    return perm[r]          # hub ranks are scattered over the id space


def make_graph(shape='wn18', seed=SEED, device=None, uniform=False, scale=1.0):
    """Returns dict(N, M, train, valid, test): int64 tensors [*, 3] in the
    reference's (s, o, p) column order, all triples distinct.  ``scale`` < 1
    shrinks the triple counts (tests)."""
    N, M, T, V, Te = SHAPES[shape] if isinstance(shape, str) else shape
    T, V, Te = int(T * scale), int(V * scale), int(Te * scale)
    device = torch.device(device) if device is not None else torch.device('cpu')
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    need = T + V + Te
    if uniform:
        ecdf = torch.linspace(1.0 / N, 1.0, N, device=device, dtype=torch.float64)
    else:
        ecdf = _zipf_sampler(N, 0.75, device)
    pcdf = _zipf_sampler(M, 1.0, device)
    eperm = torch.randperm(N, device=device, generator=gen)
    pperm = torch.randperm(M, device=device, generator=gen)
    keys = torch.zeros(0, dtype=torch.int64, device=device)
    while keys.numel() < need:
        k = int((need - keys.numel()) * 1.15) + 1024
        s, o, p = _draw(ecdf, k, gen, eperm), _draw(ecdf, k, gen, eperm), _draw(pcdf, k, gen, pperm)
        new = (p * N + s) * N + o
        keys = torch.unique(torch.cat([keys, new]))
    keys = keys[torch.randperm(keys.numel(), device=device, generator=gen)[:need]]
    o = keys % N
    s = (keys // N) % N
    p = keys // (N * N)
    tr = torch.stack([s, o, p], 1)
    return dict(N=N, M=M, train=tr[:T], valid=tr[T:T + V], test=tr[T + V:])


def init_embeddings(model, N, M, d, seed=7, device=None):
    """init_nunif values followed by the model's post-hook (row-normalised E for
    TransE, normless1 for HolE) -- fp32 tensors (E, R or W)."""
    device = torch.device(device) if device is not None else torch.device('cpu')
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))

    def nunif(*shape):
        bnd = np.sqrt(6) / np.sqrt(shape[-2] + shape[-1])
        return (torch.rand(*shape, device=device, generator=gen) * 2 - 1) * float(bnd)

    E = nunif(N, d)
    if model == 'transe':
        E = E / torch.linalg.vector_norm(E, dim=1, keepdim=True)
        return E, nunif(M, d)
    if model == 'hole':
        # trained HolE rows sit inside the unit ball (normless1); spread the norms
        E = E / torch.linalg.vector_norm(E, dim=1, keepdim=True)
        E = E * (0.5 + 0.5 * torch.rand(N, 1, device=device, generator=gen))
        return E, nunif(M, d)
    return E, nunif(M, d, d)
