#!/usr/bin/env python
"""HolE experiment (reference: skge/run_hole.py).  `python -m skge.run_hole --fin ...`"""
from . import StochasticTrainer, PairwiseStochasticTrainer, HolE
from . import activation_functions as afs
from .experiment import Experiment
from .ranking import HolEEval, FilteredRankingEval  # noqa: F401


class ExpHolE(Experiment):

    def __init__(self):
        super(ExpHolE, self).__init__()
        self.parser.add_argument('--ncomp', type=int, help='Number of latent components (dimensions)')
        self.parser.add_argument('--rparam', type=float, help='Regularization for W', default=0)
        self.parser.add_argument('--afs', type=str, default='sigmoid', help='Activation function')
        self.evaluator = HolEEval

    def setup_trainer(self, sz, sampler):
        """skge/run_hole.py:30-58."""
        model = HolE(sz, self.args.ncomp, rparam=self.args.rparam, af=afs[self.args.afs], init=self.args.init)
        if self.args.no_pairwise:
            return StochasticTrainer(model, nbatches=self.args.nb, max_epochs=self.args.me,
                                     post_epoch=[self.callback], learning_rate=self.args.lr,
                                     samplef=sampler.sample)
        return PairwiseStochasticTrainer(model, nbatches=self.args.nb, max_epochs=self.args.me,
                                         post_epoch=[self.callback], learning_rate=self.args.lr,
                                         margin=self.args.margin, samplef=sampler.sample)


if __name__ == '__main__':
    ExpHolE().run()
