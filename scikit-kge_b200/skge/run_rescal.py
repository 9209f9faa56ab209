"""RESCAL evaluator (the reference ships none; scores follow skge/rescal.py:31-35)."""
from .ranking import RESCALEval, FilteredRankingEval  # noqa: F401
