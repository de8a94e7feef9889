"""Noise models (reference: yagremcmc/statistics/noise.py:8-61)."""
from .interface import NoiseModelInterface
from .covariance import CovarianceMatrix, DiagonalCovarianceMatrix


class CentredGaussianNoise(NoiseModelInterface):

    def __init__(self, covariance):
        if not isinstance(covariance, CovarianceMatrix):
            raise ValueError("CentredGaussianNoise need to be instantiated "
                             f"with CovarianceMatrix. Got: {type(covariance)}")
        self._cov = covariance

    @property
    def covariance(self):
        return self._cov

    def induced_norm_squared(self, vector):
        return self._cov.induced_norm_squared(vector)


class AEMNoise(CentredGaussianNoise):
    """Measurement noise inflated by the variance of the model error (reference :25-61).  Descriptor
    only: the per-chain inflation  variance = scaling * errorVariance + dataVariance  with
    scaling = min(2 max(var) / max(min(var), 1e-6), 100) (heuristic) or 1 runs inside aem_mh_kernel."""

    def __init__(self, measurementNoise, useHeuristic):
        if not isinstance(measurementNoise.covariance, DiagonalCovarianceMatrix):
            raise NotImplementedError(
                "Currently, AEM is only implemented for independent measurement noise.")
        super().__init__(measurementNoise.covariance)
        self._dataNoise = measurementNoise
        self._useHeuristic = bool(useHeuristic)

    @property
    def useHeuristic(self):
        return self._useHeuristic

    @staticmethod
    def scaling_heuristic(mVar, eps=1e-6, maxScaling=100):
        import numpy as np
        minVal = max(np.min(mVar), eps)
        return min(2. * np.max(mVar) / minVal, maxScaling)
