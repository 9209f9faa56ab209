"""Import the UNMODIFIED reference (oracle/_ref, or /root/reference in the build container) under
the four shims of SURVEY.md section 8c.  Test / benchmark infrastructure.

The reference package is called ``skge`` -- like the product package -- and its modules import
each other both as ``skge.x`` and as top-level ``x`` (``from base import ...``), so it can only be
loaded in an interpreter that has NOT imported the product: bench.py's reference arm and the
subprocess behind its ``cpu_baseline`` leg.  Shims (none touches the reference's files):
  1. a stub ``trident`` module (skge/base.py:21 imports the C++ store; the evaluators never call it),
  2. the package directory on sys.path for its non-package imports (skge/run_hole.py:4),
  3. ``collections.Hashable`` (removed in Python 3.10, used by skge/util.py),
  4. ``np.Inf`` (removed in numpy 2, used by skge/base.py:973).
"""
import collections
import collections.abc
import importlib.util
import logging
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))


def find_reference():
    for root in (os.path.join(HERE, '_ref'), os.environ.get('SKGE_REFERENCE', '/root/reference')):
        if root and os.path.isfile(os.path.join(root, 'skge', 'base.py')):
            return root
    return None


def load(root=None):
    """Returns a namespace with the reference's classes: HolE, TransE, RESCAL, HolEEval, TransEEval,
    PairwiseStochasticTrainer, StochasticTrainer, AdaGrad, SGD, base (module), root (path)."""
    root = root or find_reference()
    if root is None:
        raise RuntimeError('no reference found: run `python oracle/build_ref.py` where /root/reference exists')
    mod = sys.modules.get('skge')
    if mod is not None and not os.path.realpath(getattr(mod, '__file__', '')).startswith(os.path.realpath(root)):
        raise RuntimeError('another package named skge (%s) is already imported: the reference needs its own '
                           'interpreter' % getattr(mod, '__file__', '?'))
    sys.path[:0] = [root, os.path.join(root, 'skge')]
    sys.modules.setdefault('trident', types.ModuleType('trident'))
    collections.Hashable = collections.abc.Hashable
    import numpy as np
    if not hasattr(np, 'Inf'):
        np.Inf = np.inf
    argv, sys.argv = sys.argv, ['x']
    try:
        import skge
        import skge.base as base
        from skge import HolE, TransE, RESCAL, PairwiseStochasticTrainer, StochasticTrainer
        from skge.param import AdaGrad, SGD

        def _load(name):
            spec = importlib.util.spec_from_file_location('skge_ref_' + name, os.path.join(root, 'skge', name + '.py'))
            m = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(m)
            return m
        HolEEval = _load('run_hole').HolEEval
        TransEEval = _load('run_transe').TransEEval
    finally:
        sys.argv = argv
    assert os.path.realpath(skge.__file__).startswith(os.path.realpath(root)), skge.__file__
    logging.getLogger('EX-KG').setLevel(logging.CRITICAL)
    return types.SimpleNamespace(skge=skge, base=base, HolE=HolE, TransE=TransE, RESCAL=RESCAL, HolEEval=HolEEval,
                                 TransEEval=TransEEval, PairwiseStochasticTrainer=PairwiseStochasticTrainer,
                                 StochasticTrainer=StochasticTrainer, AdaGrad=AdaGrad, SGD=SGD, root=root)
