#!/usr/bin/env python
"""RESCAL experiment.  The reference ships no run_rescal.py; this follows run_hole.py with the
logistic trainer (the only RESCAL trainer that works in the reference) and SGD."""
from . import StochasticTrainer, RESCAL
from .experiment import Experiment
from .param import SGD
from .ranking import RESCALEval, FilteredRankingEval  # noqa: F401


class ExpRESCAL(Experiment):

    def __init__(self):
        super(ExpRESCAL, self).__init__()
        self.parser.add_argument('--ncomp', type=int, help='Number of latent components (dimensions)')
        self.parser.add_argument('--rparam', type=float, help='Regularization', default=0)
        self.parser.set_defaults(no_pairwise=True)
        self.evaluator = RESCALEval

    def setup_trainer(self, sz, sampler):
        model = RESCAL(sz, self.args.ncomp, rparam=self.args.rparam, init=self.args.init)
        return StochasticTrainer(model, nbatches=self.args.nb, max_epochs=self.args.me,
                                 post_epoch=[self.callback], learning_rate=self.args.lr,
                                 samplef=sampler.sample, param_update=SGD)


if __name__ == '__main__':
    ExpRESCAL().run()
