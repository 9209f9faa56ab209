#!/usr/bin/env python
"""HolE experiment driver -- `python -m skge.run_hole --fin ... --ncomp 150 --margin 0.2`.

Command-line surface of the reference's skge/run_hole.py:22-61 (flags --ncomp, --rparam,
--afs on top of the common Experiment flags; --no-pairwise selects the logistic trainer)."""
from . import HolE, PairwiseStochasticTrainer, StochasticTrainer, activation_functions
from .experiment import Experiment
from .ranking import FilteredRankingEval, HolEEval  # noqa: F401


class ExpHolE(Experiment):
    evaluator_class = HolEEval

    def __init__(self):
        super(ExpHolE, self).__init__()
        add = self.parser.add_argument
        add('--ncomp', type=int, help='Number of latent components (dimensions)')
        add('--rparam', type=float, default=0, help='Regularization for W')
        add('--afs', type=str, default='sigmoid', help='Activation function')
        self.evaluator = self.evaluator_class

    def setup_trainer(self, sz, sampler):
        a = self.args
        model = HolE(sz, a.ncomp, rparam=a.rparam, af=activation_functions[a.afs], init=a.init)
        common = dict(nbatches=a.nb, max_epochs=a.me, learning_rate=a.lr, post_epoch=[self.callback],
                      samplef=sampler.sample)
        if a.no_pairwise:                       # logistic loss (skge/run_hole.py:40-48)
            return StochasticTrainer(model, **common)
        return PairwiseStochasticTrainer(model, margin=a.margin, **common)


if __name__ == '__main__':
    ExpHolE().run()
