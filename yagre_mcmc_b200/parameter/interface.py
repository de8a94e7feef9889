"""Parameter value types (reference: yagremcmc/parameter/interface.py:4-37).

A parameter is an immutable coefficient array plus an `evaluate()` map to the
quantity the forward model consumes.  In the batched backend the coefficient
may be one vector [d] (shared start for every chain) or a stack [nChains, d].
"""
from abc import ABC, abstractmethod


class ParameterInterface(ABC):

    @property
    @abstractmethod
    def dimension(self):
        ...

    @property
    @abstractmethod
    def coefficient(self):
        ...

    @abstractmethod
    def evaluate(self):
        ...

    @abstractmethod
    def __eq__(self, other):
        ...

    @abstractmethod
    def clone_with(self, newCoefficient):
        ...

    # --- batched-chain additions -------------------------------------------------
    #: how the device decides "proposal == state" (skip rule, metropolisHastings.py:60-61)
    equality = 'exact'
