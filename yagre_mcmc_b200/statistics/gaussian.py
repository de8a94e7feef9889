"""Gaussian law / density (reference: yagremcmc/statistics/gaussian.py:8-66).

generate_realisation() exists for host-side use (priors in set-up scripts).  Chain
proposals never come through here: the kernels draw Philox noise themselves."""
import numpy as np

from .interface import DensityInterface


class GaussianDensity(DensityInterface):
    """Unnormalised: -1/2 ||theta - m||^2_{C^-1}  (:19-24)."""

    def __init__(self, meanVector, covariance):
        self._mean = np.asarray(meanVector, dtype=np.float64)
        self._cov = covariance

    @property
    def covariance(self):
        return self._cov

    @property
    def mean(self):
        return self._mean

    def evaluate_log(self, parameter):
        x = np.asarray(parameter.coefficient, dtype=np.float64) - self._mean
        return -0.5 * self._cov.induced_norm_squared(x)

    def device_target(self):
        return dict(g_mean=self._mean, g_prec=self._cov.precision(), g_logconst=0.0)


class Gaussian:

    def __init__(self, mean, covariance):
        self._mean = mean
        self._cov = covariance
        self._density = GaussianDensity(mean.coefficient, covariance)

    @property
    def mean(self):
        return self._mean

    @property
    def covariance(self):
        return self._cov

    @property
    def density(self):
        return self._density

    def generate_realisation(self, rng=None):
        rng = np.random.default_rng() if rng is None else rng
        xi = rng.standard_normal(self._mean.dimension)
        return self._mean.clone_with(self._mean.coefficient + self._cov.apply_chol_factor(xi))
