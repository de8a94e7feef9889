"""ParameterVector (reference: yagremcmc/parameter/vector.py:5-52): exact, element-wise
equality -- the rule the kernels use for the MLDA "sub-chain did not move" skip."""
import numpy as np

from .interface import ParameterInterface


class ParameterVector(ParameterInterface):

    equality = 'exact'

    def __init__(self, coefficient):
        self.coefficient_ = coefficient
        self.coefficientType_ = type(coefficient)
        self.dim_ = int(np.shape(coefficient)[-1]) if np.ndim(coefficient) else 1

    @classmethod
    def from_coefficient(cls, coefficient):
        return cls(coefficient)

    @classmethod
    def from_value(cls, value):
        return cls(value)

    @property
    def dimension(self):
        return self.dim_

    @property
    def coefficient(self):
        return self.coefficient_

    @property
    def nChains(self):
        """None for a single vector, else the leading (chain) extent of a stacked start."""
        return None if np.ndim(self.coefficient_) < 2 else int(np.shape(self.coefficient_)[0])

    def evaluate(self):
        return self.coefficient_

    def __eq__(self, other):
        if isinstance(other, self.__class__):
            return bool(np.array_equal(self.coefficient_, other.coefficient))
        return NotImplemented

    def clone_with(self, newCoefficient):
        if not isinstance(newCoefficient, self.coefficientType_):
            raise ValueError("Trying to change coefficient type in cloning.")
        return self.__class__(newCoefficient)
