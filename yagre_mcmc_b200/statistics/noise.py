"""Noise models (reference: yagremcmc/statistics/noise.py:8-22).  AEMNoise (:25-61) belongs to
the adaptive error model, which SURVEY 8f ranks as "next"."""
from .interface import NoiseModelInterface
from .covariance import CovarianceMatrix


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
