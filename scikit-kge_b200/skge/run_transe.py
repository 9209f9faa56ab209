"""TransE evaluator (reference: skge/run_transe.py:13-29)."""
from .ranking import TransEEval, FilteredRankingEval  # noqa: F401
