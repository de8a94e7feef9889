"""ABCs of the statistics layer (reference: yagremcmc/statistics/interface.py:6-44)."""
from abc import ABC, abstractmethod


class DensityInterface(ABC):

    @abstractmethod
    def evaluate_log(self, parameter):
        ...

    # --- batched-chain addition: how a density is lowered for the kernels ------------
    def device_target(self):
        raise NotImplementedError(
            f"{type(self).__name__} has no device implementation: explicit targets must be Gaussian "
            "(GaussianTargetDensity1d / GaussianTargetDensity2d / GaussianTargetDensity); no CPU fallback")


class CovarianceOperatorInterface(ABC):

    @property
    @abstractmethod
    def dimension(self):
        ...

    @abstractmethod
    def apply_chol_factor(self, x):
        ...

    @abstractmethod
    def apply_inverse(self, x):
        ...


class NoiseModelInterface(ABC):

    @abstractmethod
    def induced_norm_squared(self, vector):
        ...


class BayesianModelInterface(ABC):

    @property
    @abstractmethod
    def likelihood(self):
        ...

    @property
    @abstractmethod
    def prior(self):
        ...
