"""ScalarParameter (reference: yagremcmc/parameter/scalar.py:5-51): a 1-d chain state whose
equality is math.isclose (rel_tol 1e-9) -- lowered to the kernels' YG_EQ_ISCLOSE rule."""
from math import isclose

import numpy as np

from .interface import ParameterInterface


class ScalarParameter(ParameterInterface):

    equality = 'isclose'

    def __init__(self, coefficient):
        if isinstance(coefficient, float):
            raise Exception("scalar parameters must be 1-dimensional array types")
        self.coefficient_ = coefficient
        self.coeffType_ = type(coefficient)

    @classmethod
    def from_coefficient(cls, coefficient):
        return cls(coefficient)

    @classmethod
    def from_value(cls, value):
        return cls(value)

    @property
    def dimension(self):
        return 1

    @property
    def coefficient(self):
        return self.coefficient_

    @property
    def nChains(self):
        return None if np.ndim(self.coefficient_) < 2 else int(np.shape(self.coefficient_)[0])

    def evaluate(self):
        return self.coefficient_

    def __eq__(self, other):
        if isinstance(other, ScalarParameter):
            return isclose(np.ravel(self.coefficient_)[0], np.ravel(other.coefficient)[0])
        return NotImplemented

    def clone_with(self, newValue):
        if not isinstance(newValue, self.coeffType_):
            raise ValueError("Trying to change coefficient type in cloning.")
        return self.__class__(newValue)
