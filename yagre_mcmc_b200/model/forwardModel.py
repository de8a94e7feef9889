"""ForwardModel wrapper (reference: yagremcmc/model/forwardModel.py:4-19)."""
from .evaluation import EvaluationStatus


class ForwardModel:

    def __init__(self, solver):
        self.solver_ = solver

    @property
    def solver(self):
        return self.solver_

    def evaluate(self, parameter):
        """Single evaluation through the solver object (device-backed for the recognised
        solvers); a failed evaluation raises like the reference (:15-19)."""
        self.solver_.interpolate(parameter)
        self.solver_.invoke()
        if self.solver_.status == EvaluationStatus.SUCCESS:
            return self.solver_.evaluation
        raise Exception("Evaluation request failed.")
