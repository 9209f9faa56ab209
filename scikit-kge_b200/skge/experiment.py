"""Experiment driver: the command-line shell of the reference's run_transe.py /
run_hole.py (skge/base.py:83-121, 204-289, 417-449, 519-731) without trident.

    python -m skge.run_transe --fin graph.npz --test-all 50 --nb 100 --me 500 \
           --margin 2.0 --lr 0.1 --ncomp 50

``--fin`` accepts
  * ``synth:<shape>[:scale]`` -- a seeded synthetic graph (wn18, fb15k, syn1m),
  * an ``.npz`` with int arrays ``train`` [, ``valid``, ``test``] of (s, o, p) rows
    [and scalars ``N``, ``M``],
  * a text file / directory of ``train.txt`` [``valid.txt``, ``test.txt``] with one
    ``s o p`` integer triple per line.
When only training triples are given, 1 % are held out for validation and 1 %
for testing, like the reference's trident.Batcher call (skge/base.py:484-489).

Flags, result-file naming, the validation/test flow of ``ranking_callback`` and
the ``--fout`` pickle follow the reference.  The subgraph modes (--subcreate /
--subtest), ``--mode lp`` and the incremental ``--incr`` training are research
features of the fork outside the hot path and are rejected with a clear error.
"""
import argparse
import logging
import os
import timeit

import numpy as np

from . import sample
from .ranking import ranking_scores

log = logging.getLogger('EX-KG')


def load_triples(fin, seed=20261018):
    """-> (train, valid, test) lists of (s, o, p) tuples and sz = (N, N, M)."""
    N = M = None
    if fin.startswith('synth:'):
        from .synth import make_graph
        parts = fin.split(':')
        g = make_graph(parts[1], scale=float(parts[2]) if len(parts) > 2 else 1.0, device='cuda')
        sets = [g[k].cpu().numpy() for k in ('train', 'valid', 'test')]
        N, M = g['N'], g['M']
    elif fin.endswith('.npz'):
        z = np.load(fin)
        sets = [np.asarray(z[k], dtype=np.int64).reshape(-1, 3) if k in z else None
                for k in ('train', 'valid', 'test')]
        N = int(z['N']) if 'N' in z else None
        M = int(z['M']) if 'M' in z else None
    else:
        if os.path.isdir(fin):
            names = [os.path.join(fin, k + '.txt') for k in ('train', 'valid', 'test')]
        else:
            names = [fin, None, None]
        sets = [np.loadtxt(n, dtype=np.int64).reshape(-1, 3) if n and os.path.exists(n) else None for n in names]
    train, valid, test = sets
    if train is None:
        raise ValueError('no training triples found in %r' % fin)
    if valid is None or test is None:
        rng = np.random.RandomState(seed % (2 ** 31))
        perm = rng.permutation(len(train))
        k = max(1, int(0.01 * len(train)))
        valid, test, train = train[perm[:k]], train[perm[k:2 * k]], train[perm[2 * k:]]
    allt = np.concatenate([train, valid, test])
    N = N or int(allt[:, :2].max()) + 1
    M = M or int(allt[:, 2].max()) + 1
    as_list = lambda a: [tuple(map(int, t)) for t in a]  # noqa: E731
    return as_list(train), as_list(valid), as_list(test), (N, N, M)


class Experiment(object):

    def __init__(self):
        p = self.parser = argparse.ArgumentParser(prog='Knowledge Graph experiment', conflict_handler='resolve')
        p.add_argument('--margin', type=float, help='Margin for loss function')
        p.add_argument('--init', type=str, default='nunif', help='Initialization method')
        p.add_argument('--lr', type=float, help='Learning rate')
        p.add_argument('--me', type=int, help='Maximum number of epochs')
        p.add_argument('--ne', type=int, help='Numer of negative examples', default=1)
        p.add_argument('--nb', type=int, help='Number of batches')
        p.add_argument('--fout', type=str, help='Path to store model and results', default=None)
        p.add_argument('--finfo', type=str, help='Path to store additional debug info', default=None)
        p.add_argument('--fgrad', type=str, help='Path to store gradient vector updates for each entity', default=None)
        p.add_argument('--fpagerank', type=str, default=None)
        p.add_argument('--fembed', type=str, help='Path to store final embeddings', default=None)
        p.add_argument('--fin', type=str, help='Path to input data', default=None)
        p.add_argument('--ftax', type=str, default=None)
        p.add_argument('--fsub', type=str, default=None)
        p.add_argument('--embed', type=str, help='Strategy to assign embeddings', default='kognac')
        p.add_argument('--test-all', type=int, help='Evaluate Test set after x epochs', default=10)
        p.add_argument('--no-pairwise', action='store_const', default=False, const=True)
        p.add_argument('--incr', type=int, default=100)
        p.add_argument('--mode', type=str, default='rank')
        p.add_argument('--sampler', type=str, default='random-mode')
        p.add_argument('--norm', type=str, default='l1', help=' Normalization (l1(default) or l2)')
        p.add_argument('--subcreate', dest='subcreate', action='store_true')
        p.add_argument('--subtest', dest='subtest', action='store_true')
        p.add_argument('--minsubsize', type=int, default=50)
        p.add_argument('--topk', type=int, default=5)
        p.add_argument('--subalgo', type=str, default='transe')
        p.add_argument('--subdistance', type=str, default='avg')
        self.neval = -1
        self.best_valid_score = -1.0
        self.exectimes = []
        self.evaluator = None

    # -- flow (skge/base.py:204-228) --------------------------------------------------
    def run(self, argv=None):
        self.args = self.parser.parse_args(argv)
        if self.args.mode != 'rank':
            raise ValueError('Unknown experiment mode (%s): only the ranking mode is part of the B200 path'
                             % self.args.mode)
        if self.args.subcreate or self.args.subtest or self.args.incr != 100:
            raise NotImplementedError('subgraph embeddings and incremental training are outside the hot path')
        self.callback = self.ranking_callback
        self.train()

    def ranking_callback(self, trn, with_eval=False):
        """skge/base.py:240-289: log the epoch, validate every --test-all epochs, test on
        improvement, pickle the best model to --fout."""
        elapsed = timeit.default_timer() - trn.epoch_start
        self.exectimes.append(elapsed)
        if self.args.no_pairwise:
            line = "[%3d] time = %ds, loss = %f" % (trn.epoch, elapsed, trn.loss)
        else:
            line = "[%3d] time = %ds, violations = %d" % (trn.epoch, elapsed, trn.nviolations)
        log.info(line)
        self.fresult.write(line + "\n")
        if (trn.epoch % self.args.test_all == 0) or with_eval:
            t0 = timeit.default_timer()
            pos_v, fpos_v = self.ev_valid.positions(trn.model)
            fmrr_valid = ranking_scores(self.fresult, pos_v, fpos_v, trn.epoch, 'VALID')
            self.fresult.write("At epoch %d , Time spent in computing positions and scores for VALIDATION dataset = %ds\n"
                               % (trn.epoch, timeit.default_timer() - t0))
            log.debug("FMRR valid = %f, best = %f" % (fmrr_valid, self.best_valid_score))
            if fmrr_valid > self.best_valid_score or trn.epoch == self.args.me:
                self.best_valid_score = fmrr_valid
                t0 = timeit.default_timer()
                pos_t, fpos_t = self.ev_test.positions(trn.model)
                ranking_scores(self.fresult, pos_t, fpos_t, trn.epoch, 'TEST')
                self.fresult.write("At epoch %d, Time spent in computing positions and scores for TEST dataset = %ds\n"
                                   % (trn.epoch, timeit.default_timer() - t0))
                if self.args.fout is not None:
                    st = {'model': trn.model, 'pos test': pos_t, 'fpos test': fpos_t, 'pos valid': pos_v,
                          'fpos valid': fpos_v, 'exectimes': self.exectimes}
                    from .base import dumps_reference
                    with open(self.args.fout, 'wb') as fout:
                        fout.write(dumps_reference(st, 2))      # loadable by the reference too
            self.fresult.flush()
        return True

    def fit_model(self, xs, ys, sz):
        """skge/base.py:417-449."""
        if self.args.sampler == 'random-mode':
            sampler = sample.RandomModeSampler(self.args.ne, [0, 1], xs, sz)
        elif self.args.sampler == 'lcwa':
            sampler = sample.LCWASampler(self.args.ne, [0, 1, 2], xs, sz)
        else:
            raise ValueError('Unknown sampler (%s)' % self.args.sampler)
        trn = self.setup_trainer(sz, sampler)
        log.info("Fitting model %s with trainer %s and parameters %s" % (
            trn.model.__class__.__name__, trn.__class__.__name__, self.args))
        trn.fit(xs, ys)
        self.callback(trn, with_eval=True)
        return trn

    def train(self):
        """skge/base.py:519-731 (the non-incremental branch)."""
        train, valid, test, sz = load_triples(self.args.fin)
        true_triples = train + test + valid
        self.ev_test = self.evaluator(test, true_triples, self.neval)
        self.ev_valid = self.evaluator(valid, true_triples, self.neval)
        dataset = self.args.fin.replace(':', '_').split('/')[-1].split('.')[0]
        outfile = "%s-%s-%s-epochs-%s-eval-%s-margin-%s.out" % (
            dataset, "full", self.args.embed, self.args.me, self.args.test_all, self.args.margin)
        self.fresult = open(outfile, "w")
        t0 = timeit.default_timer()
        trainer = self.fit_model(train, np.ones(len(train)), sz)
        line = "Time to fit model for 100%% samples (%d epochs) = %ds" % (trainer.max_epochs, timeit.default_timer() - t0)
        log.info(line)
        self.fresult.write(line + "\n")
        self.fresult.close()
        return trainer
