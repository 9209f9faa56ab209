#!/usr/bin/env python
"""Diagnostic timing of the tcgen05 ranking kernel alone (not a test, not the bench):
sweeps N, nsplit and the candidate density to separate MMA-, memory- and epilogue-bound regimes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'scikit-kge_b200')]
import torch
from skge import kernels, _ext

def run(N, Q, d, nsplit, band, reps=3):
    dev = _ext.device()
    g = torch.Generator(device=dev); g.manual_seed(0)
    E = torch.randn(N, d, device=dev, generator=g) / d ** 0.5
    Qm = torch.randn(Q, d, device=dev, generator=g)
    Ehi, Elo = kernels.pack_f16(E, None, 1024.0)
    qs = torch.full((Q,), 256.0, device=dev)
    Qhi, Qlo = kernels.pack_f16(Qm, qs, 1.0)
    f = 1024.0 * 256.0
    t = torch.zeros(Q, device=dev)
    thi = (t + band) * f
    tlo = (t - band) * f
    cnt = torch.zeros(Q, dtype=torch.int32, device=dev)
    cap = 1 << 26
    cq = torch.empty(cap, dtype=torch.int32, device=dev); ce = torch.empty(cap, dtype=torch.int32, device=dev)
    cc = torch.zeros(1, dtype=torch.int64, device=dev)
    ms = []
    for i in range(reps + 1):
        cc.zero_(); cnt.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        kernels.rank_gemm_count(Ehi, Elo, N, 0, Qhi, Qlo, Q, d, nsplit, tlo.contiguous(), thi.contiguous(), cnt, cq, ce, cc)
        b.record(); torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    best = min(ms[1:])
    flops = 2.0 * N * d * Q
    ntiles = ((N + 127) // 128) * ((Q + 127) // 128)
    sm = min(148, (Q + 127) // 128)
    waves = -(-((Q + 127) // 128) // sm)
    clk_per_tile = best * 1e-3 * 1.965e9 / (((N + 127) // 128) * waves)
    print('N=%8d Q=%6d d=%3d nsplit=%d band=%.0e : %8.3f ms  alg %7.1f TF/s  exec %7.1f TF/s  clk/tile %7.0f  cands %d'
          % (N, Q, d, nsplit, band, best, flops / best / 1e9, flops * (3 if nsplit == 3 else 1) / best / 1e9,
             clk_per_tile, int(cc.item())), flush=True)

def probe_error(N=1000000, Q=256, d=256, nsplit=3):
    """Lower bound on the coarse pass's error near a threshold T: with thr_lo = thr_hi = T
    the kernel returns #{coarse score > T}; every disagreement with the exact fp64 count needs
    at least one pair whose error exceeds its distance to T.  Reported relative to |q| * max|e|."""
    dev = _ext.device()
    g = torch.Generator(device=dev); g.manual_seed(1)
    E = torch.randn(N, d, device=dev, generator=g) / d ** 0.5
    E = E * (0.5 + 0.5 * torch.rand(N, 1, device=dev, generator=g))
    Qm = torch.randn(Q, d, device=dev, generator=g) * 0.1
    emax = float(E.abs().max()); import math
    escale = 2.0 ** (12 - math.ceil(math.log2(emax)))
    Ehi, Elo = kernels.pack_f16(E, None, escale)
    qmax = Qm.abs().amax(dim=1)
    qs = torch.exp2(12 - torch.ceil(torch.log2(qmax)))
    Qhi, Qlo = kernels.pack_f16(Qm, qs.contiguous(), 1.0)
    S = (E.double() @ Qm.double().t()).t().contiguous()      # [Q, N] exact
    qn = torch.linalg.vector_norm(Qm.double(), dim=1)
    en = float(torch.linalg.vector_norm(E.double(), dim=1).max())
    sig = S.std(dim=1)
    cap = 1 << 20
    cq = torch.empty(cap, dtype=torch.int32, device=dev); ce = torch.empty(cap, dtype=torch.int32, device=dev)
    worst = 0.0
    for k in (0.0, 0.5, 1.0, 2.0, -1.0):
        T = k * sig                                            # per-query threshold
        f = (qs.double() * escale)
        thr = (T * f).float().contiguous()
        Tq = thr.double() / f                                  # the threshold the kernel really used
        cnt = torch.zeros(Q, dtype=torch.int32, device=dev)
        cc = torch.zeros(1, dtype=torch.int64, device=dev)
        kernels.rank_gemm_count(Ehi, Elo, N, 0, Qhi, Qlo, Q, d, nsplit, thr, thr, cnt, cq, ce, cc)
        torch.cuda.synchronize()
        exact = (S > Tq[:, None]).sum(dim=1)
        diff = (cnt.long() - exact)
        rel = torch.zeros(Q, dtype=torch.float64, device=dev)
        for qi in torch.nonzero(diff).flatten().tolist():
            dq = int(diff[qi])
            row = S[qi] - Tq[qi]
            side = (-row[row <= 0]) if dq > 0 else row[row > 0]
            kth = torch.kthvalue(side, abs(dq)).values
            rel[qi] = kth / (qn[qi] * en)
        worst = max(worst, float(rel.max()))
        print('T=%+.1f sigma: queries with a count mismatch %3d / %d, max |diff| %d, implied error >= %.3g (2^%.1f) of |q||e|max'
              % (k, int((diff != 0).sum()), Q, int(diff.abs().max()), float(rel.max()),
                 math.log2(float(rel.max())) if float(rel.max()) > 0 else -99), flush=True)
    print('nsplit=%d worst implied relative error: %.3g = 2^%.2f' % (nsplit, worst, math.log2(worst) if worst > 0 else -99))


if __name__ == '__main__':
    if 'probe' in sys.argv:
        probe_error(nsplit=3)
        probe_error(nsplit=1)
        sys.exit(0)
    Q = 148 * 128
    for N in (50000, 200000, 1000000):
        for ns in (1, 3):
            run(N, Q, 256, ns, 0.0)
    run(1000000, Q, 256, 3, 3e-5)
    run(1000000, Q, 256, 3, 3e-4)
    run(1000000, Q, 128, 3, 0.0)
    run(1000000, Q, 64, 3, 0.0)
    run(1000000, 2 * Q, 256, 3, 0.0)
