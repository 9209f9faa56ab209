"""skge -- B200-native drop-in for the scikit-kge training / link-prediction hot path.

Export list identical to the reference's skge/__init__.py:1-5.
"""
from .hole import HolE
from .rescal import RESCAL
from .transe import TransE
from .base import StochasticTrainer, PairwiseStochasticTrainer
from .actfun import afuns as activation_functions
from .version import __version__
