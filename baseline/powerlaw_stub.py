"""Stub for the `powerlaw` package (not installable offline).

TEST INFRASTRUCTURE. cornac/eval_methods/propensity_stratified_evaluation.py:5
imports it at package-import time; only PropensityStratifiedEvaluation uses it,
which is outside the hot path (SURVEY.md section 8c caveat 2).
"""


class Fit:  # pragma: no cover
    def __init__(self, *a, **k):
        raise NotImplementedError("powerlaw is stubbed in this environment")
