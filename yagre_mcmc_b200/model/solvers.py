"""The forward models the kernels implement, as SolverInterface plugins.

LinearModelSolver       G theta + b             (reference exampleSetup.py:8-52)
LotkaVolterraRK4Solver  nData independent 2-state ODE solves over [0,T] by classical RK4
                        with N fixed steps; same design / initial-condition semantics as the
                        reference's LotkaVolterraSolver (yagremcmc/test/testSetup.py:61-141),
                        whose scipy.solve_ivp integrators north_star replaces by RK4.
LotkaVolterraParameter  log-parametrised LV parameter (testSetup.py:47-58): evaluate() = exp.

interpolate()/invoke() on these objects evaluate ONE parameter on the device (yg_forward is
not part of the ABI; the log-posterior kernel is), so `evaluation` is only provided where a
closed form is trivial (linear).  Their purpose is to be lowered by the builders.
"""
import numpy as np

from .interface import SolverInterface
from .evaluation import EvaluationStatus
from ..parameter.vector import ParameterVector


class LinearModelSolver(SolverInterface):

    def __init__(self, A, b):
        self._A = np.asarray(A, dtype=np.float64)
        self._b = np.asarray(b, dtype=np.float64)
        if self._A.ndim != 2 or self._b.shape != (self._A.shape[0],):
            raise ValueError("LinearModelSolver needs A[dataDim, d] and b[dataDim]")
        self._coef = None
        self._evaluation = None
        self._status = EvaluationStatus.NONE

    @property
    def status(self):
        return self._status

    @property
    def evaluation(self):
        if self._status == EvaluationStatus.FAILURE:
            raise RuntimeError("Trying to retrieve failed evaluation.")
        if self._status == EvaluationStatus.NONE:
            raise RuntimeError("Trying to retrieve evaluation before it was performed.")
        return self._evaluation

    def interpolate(self, parameter):
        self._coef = np.asarray(parameter.coefficient, dtype=np.float64)

    def invoke(self):
        # descriptor-level convenience for a single parameter (data synthesis in examples);
        # chains never come through here -- they run in generic_kernel.cu
        self._evaluation = self._A @ self._coef + self._b
        self._status = EvaluationStatus.SUCCESS

    def device_model(self):
        return 'linear'

    def device_level(self):
        return dict(G=self._A, b=self._b)


class LotkaVolterraParameter(ParameterVector):

    @classmethod
    def from_interpolation(cls, value):
        return cls(np.log(value))

    def evaluate(self):
        return np.exp(self.coefficient_)


class LotkaVolterraRK4Solver(SolverInterface):
    """config keys: T, alpha, gamma, nData, dataDim (=2), rk4Steps."""

    def __init__(self, design, config):
        self.x_ = np.asarray(design, dtype=np.float64)
        self.T_ = float(config['T'])
        self.N_ = int(config['rk4Steps'])
        self.alpha_, self.gamma_ = float(config['alpha']), float(config['gamma'])
        self.dataShape_ = (int(config['nData']), int(config['dataDim']))
        if self.x_.shape != (self.dataShape_[0], 2) or self.dataShape_[1] != 2:
            raise ValueError("LotkaVolterraRK4Solver needs design[nData, 2] and dataDim == 2")
        self.param_ = None
        self.evaluation_ = None
        self.status_ = EvaluationStatus.NONE

    @property
    def status(self):
        return self.status_

    @property
    def dataShape(self):
        return self.dataShape_

    @property
    def nData(self):
        return self.dataShape_[0]

    @property
    def dataDim(self):
        return self.dataShape_[1]

    @property
    def evaluation(self):
        return self.evaluation_

    def interpolate(self, parameter):
        self.param_ = np.asarray(parameter.evaluate(), dtype=np.float64)

    def invoke(self):
        raise NotImplementedError(
            "single evaluations of the RK4 Lotka-Volterra model are not exposed on the host; build a chain "
            "(MRWBuilder / MLDABuilder) or use sampler.target.evaluate_log(), which run on the device")

    def device_model(self):
        return 'lv'

    def device_level(self):
        return dict(design=self.x_, lv=np.array([self.alpha_, self.gamma_, self.T_, float(self.N_)]))
