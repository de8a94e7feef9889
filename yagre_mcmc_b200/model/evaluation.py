"""Evaluation status of a forward solve (reference: yagremcmc/model/evaluation.py:5-9)."""
from enum import Enum, unique


@unique
class EvaluationStatus(Enum):
    NONE = -1
    SUCCESS = 0
    FAILURE = 1
