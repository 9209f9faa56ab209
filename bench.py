#!/usr/bin/env python
"""Benchmark of the scikit-kge hot path on B200 (contract: see DESIGN.md section "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One JSON line on rank 0.  Headline metric: filtered-rank queries/s on BASELINE.json's
config 5 (synthetic 1M entities / 1k relations, HolE d=256; 2 x 100k test queries,
raw + filtered ranks, both directions), entity table sharded over the N GPUs
(strong scaling: the graph and the query set are fixed).  A "step" is one full
ranking pass over the test set.  At N=1 the line also carries `train` (HolE
d=256 minibatch steps of 500k positives, config 5's batch size), `roofline`
(the coarse contraction kernel), `cpu_baseline` (the oracle port on host cores)
and `e2e` (through FilteredRankingEval.positions with host inputs / outputs).

--impl reference times the reference's CPU algorithm (the numpy oracle port --
the reference is Python and cannot travel to the GPU box) on a bounded sample
of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, 'scikit-kge_b200')):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (graph shape, model, d, description)
    'cfg5': ('syn1m', 'hole', 256, 'cfg5: synthetic 1M entities / 1k relations / 50M triples, HolE d=256, '
                                    'filtered ranking of 100k test triples (200k queries)'),
    'cfg2': ('wn18', 'hole', 150, 'cfg2: WN18 shape (40,943 ent / 18 rel), HolE d=150, filtered ranking of 5k test triples'),
    'cfg4': ('fb15k', 'transe', 200, 'cfg4: FB15k shape (14,951 ent / 1,345 rel), TransE d=200 L1, filtered ranking of 59,071 test triples'),
    'cfg1': ('wn18', 'transe', 50, 'cfg1: WN18 shape, TransE d=50 L1, filtered ranking of 5k test triples'),
}


# per-launch DRAM bytes (read + write) and tensor-pipe active % of the coarse kernel at the full
# config-5 size on one GPU, from the committed ncu --set full captures
NCU_GEMM = {'tcgen05-f16x3': (12.875e9 + 0.180e9, 69.2), 'tcgen05-f16x2': (13.641e9 + 0.178e9, 53.5)}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='cfg5', choices=sorted(WORKLOADS))
    ap.add_argument('--test-triples', type=int, default=0, help='override the number of test triples')
    ap.add_argument('--true-triples', type=int, default=0, help='override the number of known-true triples used for filtering')
    ap.add_argument('--no-train', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--train-batches', type=int, default=6)
    ap.add_argument('--engine', default='auto', choices=['auto', 'sweep', 'umma'])
    ap.add_argument('--nsplit', type=int, default=0, choices=[0, 1, 2, 3],
                    help='tcgen05 engine: fp16 products on the tensor cores (0 = the evaluator default)')
    return ap.parse_args()


def peaks():
    f = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(f):
        p = json.load(open(f))
        return dict(hbm=p['hbm_gbs'], tf_burst=p['bf16_tflops'], tf_sust=p['bf16_tflops_sustained'], src='measured')
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src='fallback')


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(gpu_index), '--query-gpu=' + self.Q,
                                       '--format=csv,noheader,nounits', '-lms', '100'],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            pass

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [l.split(', ') for l in open(self.f.name).read().strip().splitlines() if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.strip().lower().startswith('active'):
                        reasons.add(nm)
            except Exception:
                continue
        if sm:
            out = dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm))
        return out


# ---------------------------------------------------------------------------
# reference arm: the numpy oracle port on the host cores
# ---------------------------------------------------------------------------

def cpu_rank_sample(model, E, R, N, M, d, queries_per_relation, nq=2, seed=0):
    """One 'reference step': prepare() for one relation + nq queries of that relation
    (skge/base.py:937-1017), float64 numpy.  Returns (t_prepare, t_per_query)."""
    from oracle import cpu_oracle as orc
    rng = np.random.default_rng(seed)
    p = int(rng.integers(M))
    prepare, scores_o, scores_s = orc._eval_hooks(model, E, R)
    t0 = time.perf_counter()
    prepare(p)
    t1 = time.perf_counter()
    for i in range(nq):
        s, o = int(rng.integers(N)), int(rng.integers(N))
        sc = (scores_o(s, p) if i % 2 == 0 else scores_s(o, p)).flatten()
        tgt = o if i % 2 == 0 else s
        order = np.argsort(sc)[::-1]
        _ = int(np.where(order == tgt)[0][0]) + 1
        sc[rng.integers(N, size=8)] = -np.inf           # the filter list of the query
        order = np.argsort(sc)[::-1]
        _ = int(np.where(order == tgt)[0][0]) + 1
    t2 = time.perf_counter()
    return t1 - t0, (t2 - t1) / nq


def host_tables(model, N, M, d):
    rng = np.random.default_rng(7)
    E = rng.uniform(-1, 1, (N, d)) * (np.sqrt(6) / np.sqrt(N + d))
    E /= np.linalg.norm(E, axis=1, keepdims=True)
    R = rng.uniform(-1, 1, (M, d)) * (np.sqrt(6) / np.sqrt(M + d))
    return E, R


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from skge.synth import SHAPES
    shape, model, d, desc = WORKLOADS[args.workload]
    N, M, T, V, Te = SHAPES[shape]
    te = args.test_triples or Te
    qpr = max(1.0, 2.0 * te / min(M, te))       # queries that share one prepare() in the full workload
    E, R = host_tables(model, N, M, d)
    tp, tq = [], []
    for i in range(args.warmup):
        cpu_rank_sample(model, E, R, N, M, d, qpr, seed=100 + i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        a, b = cpu_rank_sample(model, E, R, N, M, d, qpr, seed=i)
        tp.append(a)
        tq.append(b)
    wall = time.perf_counter() - t0
    tpm, tqm = float(np.mean(tp)), float(np.mean(tq))
    value = qpr / (tpm + qpr * tqm)
    cores = len(os.sched_getaffinity(0))
    sample = ('per step: prepare() of one relation + 2 queries (raw+filtered argsort ranks) on the full N=%d table; '
              'value extrapolates to the workload\'s %.0f queries per relation: q/(t_prepare + q*t_query), '
              't_prepare=%.2fs t_query=%.3fs' % (N, qpr, tpm, tqm))
    line = {
        'impl': 'reference', 'metric': 'filtered-rank queries/s', 'value': value, 'unit': 'queries/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1000.0 * wall / max(1, args.steps), 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': desc, 'test_triples': te, 'engine': 'numpy oracle port (oracle/cpu_oracle.py)'},
        'cpu_baseline': {'value': value, 'unit': 'queries/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'queries/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------

def run_b200(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)

    import skge
    from skge import kernels, ranking
    from skge.synth import make_graph, init_embeddings, SHAPES
    shape, model, d, desc = WORKLOADS[args.workload]
    N, M, T, V, Te = SHAPES[shape]
    te = args.test_triples or Te

    # --- synthetic inputs (identical on every rank: same seed) -------------------
    torch.manual_seed(1234)
    t_setup = time.perf_counter()
    ntrue = args.true_triples or (T + V + Te)
    g = make_graph((N, M, max(ntrue - V - te, 1), V, te), device=dev)
    true = torch.cat([g['train'], g['valid'], g['test']])
    test = g['test']
    cls = {'hole': skge.HolE, 'transe': skge.TransE}[model]
    mdl = cls((N, N, M), d)
    E, R = init_embeddings(model, N, M, d, device=dev)
    mdl.E.data.copy_(E)
    mdl.R.data.copy_(R)
    del E, R
    Ev = {'hole': ranking.HolEEval, 'transe': ranking.TransEEval}[model]
    ev = Ev(test, true)
    if args.engine != 'auto':
        ev.engine = args.engine
    if args.nsplit:
        ev.nsplit = args.nsplit
    del true, g
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t_setup
    Q = 2 * te

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # --- device-resident metric ("value") ----------------------------------------
    ev._device_state()                       # filter index + query descriptors resident
    clocks = ClockSampler(local)             # sampled through warm-up + timed region (same load)
    for _ in range(max(3, args.warmup)):
        ev.count_pass(mdl)
    ranking._SweepEngine.timing = True
    ranking.TIMINGS.clear()
    l0 = kernels.LAUNCHES['n']
    ms = timed(lambda: ev.count_pass(mdl), args.steps)
    clk = clocks.stop()
    launches = kernels.LAUNCHES['n'] - l0
    ranking._SweepEngine.timing = False
    kt = [(a.elapsed_time(b), w) for a, b, w in ranking.TIMINGS]
    ranking.TIMINGS.clear()
    stats = dict(ev.last_stats)
    value = Q * args.steps / (ms / 1000.0)

    # --- end to end through the public API (host in, host out) --------------------
    def e2e_step():
        pos, fpos = ev.positions(mdl)
        return pos, fpos
    e2e_step()
    wall0 = time.perf_counter()
    barrier()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - wall0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e = Q * args.steps / float(e2e_s.item())

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    pk = peaks()
    # roofline of the dominant kernel (the coarse query x entity contraction)
    tensor = model != 'transe'
    if kt:
        avg_ms = float(np.mean([t for t, _ in kt]))
        work = float(np.mean([w for _, w in kt]))
        if tensor:
            achieved = work / (avg_ms * 1e-3) / 1e12
            peak = pk['tf_sust']
            traffic = pipe_pct = None
            eng = str(stats.get('engine', ''))
            if args.workload == 'cfg5' and world == 1 and te == Te:
                # dram read + write of one launch and tensor-pipe activity from ncu --set full
                # (profiles/r01_rank_cfg5_tcgen05_v2.txt, profiles/r01_rank_cfg5_tcgen05_refine.txt)
                traffic, pipe_pct = NCU_GEMM.get(eng, (None, None))
            nprod = {'tcgen05-f16x3': 3, 'tcgen05-f16x2': 2}.get(eng, 1)
            roof = {'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s',
                    'frac': achieved / peak, 'traffic': traffic, 'kernel': eng,
                    'mma_products_per_algorithmic_mac': nprod,
                    'executed_tflops': achieved * nprod,
                    'tensor_pipe_active_pct_ncu': pipe_pct,
                    'ncu_note': ('captured one commit earlier (fp16 e_lo rows in the refinement; the 8-bit rows '
                                 'gather half of those bytes)') if nprod == 2 and traffic else None,
                    'launch_ms': avg_ms, 'launches_timed': len(kt), 'peak_source': pk['src'] + ' bf16 sustained',
                    'algorithmic_flops_per_launch': work,
                    'kernel_share_of_step': sum(t for t, _ in kt) / max(ms, 1e-9)}
        else:
            n_shard = stats['shard'][1] - stats['shard'][0]
            # reference access pattern: one sweep of the fp32 table per query (4*N*d bytes)
            bytes_per_launch = work / 2.0 * 4.0
            achieved = bytes_per_launch / (avg_ms * 1e-3) / 1e9
            roof = {'bound': 'hbm', 'achieved': achieved, 'peak': pk['hbm'], 'unit': 'GB/s',
                    'frac': achieved / pk['hbm'], 'traffic': None, 'kernel': stats.get('engine'),
                    'launch_ms': avg_ms, 'launches_timed': len(kt), 'peak_source': pk['src'],
                    'note': 'algorithmic bytes = 4*N*d per query (the reference re-reads the table per query); '
                            'query tiling makes the kernel FP32-ALU bound, so frac > 1 is expected',
                    'kernel_share_of_step': sum(t for t, _ in kt) / max(ms, 1e-9)}
    else:
        roof = None

    line = {
        'metric': 'filtered-rank queries/s', 'value': value, 'unit': 'queries/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': max(3, args.warmup), 'ms_per_step': ms / args.steps,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': stats.get('dtype', 'f32+f64'), 'data': 'synthetic',
        'config': {'workload': desc, 'queries': Q, 'entities': N, 'relations': M, 'd': d,
                   'filter_pairs': stats.get('filter_pairs'), 'band_candidates_last_step': stats.get('candidates'),
                   'sharding': 'entity rows / %d ranks' % world, 'l2': 'inputs larger than L2 (entity table %.0f MB)'
                   % (N * d * 4 / 1e6) if N * d * 4 > 126e6 else 'entity table is L2-resident (%.1f MB); flagged' % (N * d * 4 / 1e6),
                   'engine': stats.get('engine'), 'setup_s': setup_s},
        'clocks': clk,
        'e2e': {'value': e2e, 'unit': 'queries/s', 'h2d_bytes_per_step': ev.h2d_bytes(),
                'd2h_bytes_per_step': 2 * Q * 4, 'api': '%s(test, true).positions(model)' % Ev.__name__},
        'gpu_launches': launches,
        'roofline': roof,
    }

    # --- training throughput on 1 GPU (config 5's batch size) ----------------------
    if world == 1 and not args.no_train:
        try:
            line['train'] = bench_train(args, mdl, N, M, d, model, pk)
        except Exception as e:  # keep the headline line even if the side benchmark fails
            line['train'] = {'error': repr(e)}

    # --- CPU baseline (oracle port) --------------------------------------------------
    if world == 1 and not args.no_cpu:
        try:
            qpr = max(1.0, 2.0 * te / min(M, te))
            Eh = mdl.E.data.double().cpu().numpy()
            Rh = mdl.R.data.double().cpu().numpy()
            tp, tq = cpu_rank_sample(model, Eh, Rh, N, M, d, qpr, nq=4)
            v = qpr / (tp + qpr * tq)
            line['cpu_baseline'] = {
                'value': v, 'unit': 'queries/s', 'cores': len(os.sched_getaffinity(0)), 'kind': 'port',
                'sample': 'prepare() of one relation + 4 queries on the full table, extrapolated to %.0f queries per '
                          'relation (t_prepare=%.2fs, t_query=%.3fs); numpy float64, default threading' % (qpr, tp, tq)}
        except Exception as e:
            line['cpu_baseline'] = {'error': repr(e)}

    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def bench_train(args, mdl, N, M, d, model, pk):
    """Fused minibatch steps at config-5 batch size through PairwiseStochasticTrainer.fit."""
    import torch
    import skge
    from skge import kernels
    from skge.param import AdaGrad
    from skge.sample import RandomModeSampler
    from skge.synth import make_graph
    dev = mdl.E.data.device
    B = 500000 if N >= 1000000 else None
    nb = args.train_batches
    T = (B * nb) if B else None
    if T is None:
        from skge.synth import SHAPES
        T = [v for k, v in SHAPES.items() if v[0] == N][0][2]
        nb = 100
    g = make_graph((N, M, T, 1, 1), device=dev, seed=99)
    xs = g['train'].cpu().numpy()
    smp = RandomModeSampler(1, [0, 1], xs, (N, N, M))
    margin = 0.2 if model == 'hole' else 2.0
    times = []

    def cb(trn):
        torch.cuda.synchronize()
        times.append(time.perf_counter() - trn.epoch_start)
        return True

    trn = skge.PairwiseStochasticTrainer(mdl, nbatches=nb, margin=margin, max_epochs=5, learning_rate=0.1,
                                         samplef=smp.sample, param_update=AdaGrad, post_epoch=[cb])
    mdl.track_counters = False
    l0 = kernels.LAUNCHES['n']
    trn.fit(xs, np.ones(len(xs), dtype=np.float32))
    launches = kernels.LAUNCHES['n'] - l0
    best = float(np.median(times[1:]))   # violations (hence work) shrink as the model learns: report the median epoch
    P = 2 * (T // nb)
    # algorithmic bytes of one minibatch: 4*d*[4P + c*(U_E+U_R)] + 24P, c = 4 (AdaGrad),
    # with the unique-row counts of the last minibatch
    nv, ue, ur, _ = trn._counts.tolist()
    bytes_batch = 4.0 * d * (4 * P + 4 * (ue + ur)) + 24.0 * P
    ach = bytes_batch / (best / nb) / 1e9
    # worst-case algorithmic bytes per pair: 4*d*[4P + c(U_E+U_R)], U <= 4P, c = 4
    return {'metric': 'train triples/s', 'value': T / best, 'unit': 'triples/s', 'epoch_s': best,
            'epochs_timed': len(times) - 1, 'model': model, 'd': d, 'batch_positives': T // nb,
            'pairs_per_batch': P, 'nbatches': nb, 'triples': T, 'violations_last_epoch': trn.nviolations,
            'gpu_launches': launches, 'sampler': 'on-device RandomModeSampler(1, [0,1])', 'optimizer': 'AdaGrad',
            'epoch_times_s': [round(t, 5) for t in times],
            'roofline': {'bound': 'hbm', 'achieved': ach, 'peak': pk['hbm'], 'unit': 'GB/s', 'frac': ach / pk['hbm'],
                         'algorithmic_bytes_per_minibatch': bytes_batch, 'minibatch_ms': 1e3 * best / nb,
                         'unique_rows_last_minibatch': [ue, ur], 'violating_pairs_last_minibatch': nv,
                         'peak_source': pk['src'], 'traffic': None,
                         'note': 'whole fused step (sampler + pair kernel + sort + segmented update), wall clock per minibatch'},
            'note': 'steady-state epochs through PairwiseStochasticTrainer.fit (first epoch excluded)'}


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_b200(a)
