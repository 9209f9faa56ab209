#!/usr/bin/env python
"""TransE experiment (reference: skge/run_transe.py).  `python -m skge.run_transe --fin ...`"""
from . import TransE, PairwiseStochasticTrainer
from .experiment import Experiment
from .ranking import TransEEval, FilteredRankingEval  # noqa: F401


class ExpTransE(Experiment):

    def __init__(self):
        super(ExpTransE, self).__init__()
        self.parser.add_argument('--ncomp', type=int, help='Number of latent components (dimensions)')
        self.evaluator = TransEEval

    def setup_trainer(self, sz, sampler):
        """skge/run_transe.py:39-65 (no param_update given: AdaGrad, skge/base.py:1216)."""
        model = TransE(sz, self.args.ncomp, l1=self.args.norm == 'l1', init=self.args.init)
        return PairwiseStochasticTrainer(
            model, nbatches=self.args.nb, margin=self.args.margin, max_epochs=self.args.me,
            learning_rate=self.args.lr, samplef=sampler.sample, post_epoch=[self.callback],
            file_grad=self.args.fgrad, file_embed=self.args.fembed)


if __name__ == '__main__':
    ExpTransE().run()
