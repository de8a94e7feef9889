"""Bayesian regression model = likelihood + prior (reference: yagremcmc/statistics/bayesModel.py:5-29)."""
from .interface import BayesianModelInterface
from ..utility.hierarchy import Hierarchy, SharedComponent, HierarchyBase


class BayesianRegressionModel(BayesianModelInterface):

    def __init__(self, likelihood, prior):
        self._likelihood = self._single(likelihood, "likelihood")
        self._prior = self._single(prior, "prior")

    @staticmethod
    def _single(component, name):
        if isinstance(component, HierarchyBase):
            if isinstance(component, SharedComponent):
                return component.level(0)
            raise RuntimeError(f"Setting non-hierarchical {name} with a {name} hierarchy")
        return component

    @property
    def prior(self):
        return self._prior

    @property
    def likelihood(self):
        return self._likelihood
