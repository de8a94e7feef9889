"""Forward-model plugin boundary (reference: yagremcmc/model/interface.py:7-67).

In the reference a solver is arbitrary Python called once per likelihood evaluation.
A Python callable cannot run inside a CUDA kernel, so the batched backend accepts only
solvers that carry a *device descriptor*: `device_model()` returns the name of a model
the kernels implement and `device_level()` the plain arrays of one hierarchy level.
Anything else is refused with NotImplementedError at build time -- there is no CPU
fallback (north_star).
"""
from abc import ABC, abstractmethod


class SolverInterface(ABC):

    @property
    @abstractmethod
    def status(self):
        ...

    @property
    @abstractmethod
    def evaluation(self):
        ...

    @abstractmethod
    def interpolate(self, parameter):
        ...

    @abstractmethod
    def invoke(self):
        ...

    # --- batched-chain additions -------------------------------------------------
    def device_model(self):
        raise NotImplementedError(
            f"{type(self).__name__} has no device implementation: the B200 backend runs the linear model "
            "(LinearModelSolver) and the fixed-step RK4 Lotka-Volterra model (LotkaVolterraRK4Solver) only")

    def device_level(self):
        raise NotImplementedError
