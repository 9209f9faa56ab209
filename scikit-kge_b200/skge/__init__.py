"""skge -- B200-native drop-in for the scikit-kge training / link-prediction hot path.

The names exported here are the ones user code imports from the reference package
(its skge/__init__.py:1-5): the three model classes, the two trainers and the
activation-function registry.  Everything they compute runs in libskge_b200.so
(hand-written sm_100a kernels, see include/skge_b200.h); there is no CPU fallback.
"""
from .version import __version__
from .actfun import afuns as activation_functions
from .base import PairwiseStochasticTrainer, StochasticTrainer
from .transe import TransE
from .rescal import RESCAL
from .hole import HolE

__all__ = ['HolE', 'RESCAL', 'TransE', 'StochasticTrainer', 'PairwiseStochasticTrainer',
           'activation_functions', '__version__']
