"""Hierarchy of Bayesian regression models (reference: yagremcmc/statistics/modelHierarchy.py:6-53)."""
from ..utility.hierarchy import HierarchyBase, Hierarchy
from .bayesModel import BayesianRegressionModel


class BayesianRegressionModelHierarchy(Hierarchy):

    def __init__(self, likelihood, prior):
        parts = [("prior", prior), ("likelihood", likelihood)]
        for name, inst in parts:
            if not isinstance(inst, HierarchyBase):
                raise ValueError(f"Argument '{name}' must be derived from the Hierarchy base class. "
                                 f"Received type: {type(inst).__name__}.")
        if len({inst.size for _, inst in parts}) > 1:
            report = "\n".join(f" - {name}: size {inst.size}" for name, inst in parts)
            raise ValueError("Hierarchies have mismatched sizes. The following mismatches were found:\n" + report)
        super().__init__([BayesianRegressionModel(likelihood.level(l), prior.level(l))
                          for l in range(likelihood.size)])
