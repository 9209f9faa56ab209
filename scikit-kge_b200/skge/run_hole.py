"""HolE evaluator (reference: skge/run_hole.py:10-19)."""
from .ranking import HolEEval, FilteredRankingEval  # noqa: F401
