#!/usr/bin/env python
"""Diagnostic: what does ONE rank of an N-way sharded config-5 ranking pass cost on its own GPU?
Emulates shard 0 of `world` on a single GPU (count_pass(world=(0, world)): same kernels, no
all-reduce) and prints the pass time next to the ideal T(world=1) / world, plus the time of the
phases that do not shrink with the shard (query construction, packing, filter-pair selection)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'scikit-kge_b200')]
import torch
import skge
from skge import kernels, ranking
from skge.synth import make_graph, init_embeddings, SHAPES


def ev_ms(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main(worlds=(1, 2, 4, 8)):
    dev = torch.device('cuda')
    N, M, T, V, Te = SHAPES['syn1m']
    g = make_graph((N, M, T, V, Te), device=dev)
    true = torch.cat([g['train'], g['valid'], g['test']])
    mdl = skge.HolE((N, N, M), 256)
    E, R = init_embeddings('hole', N, M, 256, device=dev)
    mdl.E.data.copy_(E)
    mdl.R.data.copy_(R)
    ev = ranking.HolEEval(g['test'], true)
    ev.nsplit = int(os.environ.get('SKGE_NSPLIT', ev.nsplit))
    print('nsplit', ev.nsplit, flush=True)
    del true, g
    st = ev._device_state()
    Q = 2 * Te
    if os.environ.get('SKGE_VERIFY'):
        # full-size check: the three engines settle their bands in fp64, so their counts must be equal
        ref = None
        for eng, ns in (('umma', 3), ('umma', 2), ('sweep', 0)):
            ev.engine, ev.nsplit = eng, ns
            t0 = time.perf_counter()
            c = ev.count_pass(mdl, world=(0, 1))
            torch.cuda.synchronize()
            print('%s nsplit %d: %.0f ms, candidates %d, sum of counts %d / %d' % (
                ev.last_stats['engine'], ns, (time.perf_counter() - t0) * 1e3, ev.last_stats['candidates'],
                int(c[0].sum().item()), int(c[1].sum().item())), flush=True)
            assert ref is None or torch.equal(c, ref), 'engines disagree'
            ref = c if ref is None else ref
        print('verified: tcgen05 x3, tcgen05 x2 (refine) and the fp32 sweep give identical counts on all %d queries' % Q)
        ev.engine = 'auto'
        ev.nsplit = int(os.environ.get('SKGE_NSPLIT', 0))
    t1 = None
    for w in worlds:
        ms = ev_ms(lambda: ev.count_pass(mdl, world=(0, w)))
        t1 = t1 or ms
        print('world %d: shard-0 pass %.1f ms (ideal %.1f ms) -> %.2f M queries/s if every rank takes this long'
              % (w, ms, t1 / w, Q / ms / 1e3), flush=True)
    # phase breakdown of one world-1 pass (CUDA events around the library calls)
    spans = {}

    def wrap(name):
        fn = getattr(kernels, name)

        def timed(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a, **k)
            e1.record()
            spans.setdefault(name, []).append((e0, e1))
            return r
        setattr(kernels, name, timed)
        return fn
    names = ['make_queries', 'rank_rescore', 'rank_gemm_count', 'pack_f16', 'query_scale', 'rank_sweep']
    saved = {n: wrap(n) for n in names if hasattr(kernels, n)}
    bw = int(os.environ.get('SKGE_BREAKDOWN_WORLD', '1'))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ev.count_pass(mdl, world=(0, bw))
    torch.cuda.synchronize()
    print('breakdown pass (world %d): %.1f ms wall' % (bw, (time.perf_counter() - t0) * 1e3))
    for n, fn in saved.items():
        setattr(kernels, n, fn)
    print('one pass: ' + ', '.join('%s %.2f ms x%d' % (n, sum(a.elapsed_time(b) for a, b in v), len(v))
                                   for n, v in spans.items()), flush=True)
    enorm = float(torch.linalg.vector_norm(mdl.E.data, dim=1).max().item())
    mq = ev_ms(lambda: kernels.make_queries(ev.model_code, mdl.E.data, mdl.R.data, st['kind'], st['given'], st['rel'],
                                            st['target'], enorm, 2.0 ** -17))
    nm = ev_ms(lambda: torch.linalg.vector_norm(mdl.E.data, dim=1).max().item())
    print('fixed per rank: make_queries(%d) %.2f ms, max row norm of E %.2f ms' % (Q, mq, nm), flush=True)


if __name__ == '__main__':
    main()
