"""Explicit Gaussian target densities (reference fixtures: yagremcmc/test/testSetup.py:15-44).

GaussianTargetDensity1d  -1/2 (m - theta)^2 / var                      (:15-29)
GaussianTargetDensity2d  scipy.stats.multivariate_normal(m, C).logpdf  (:32-44); the
                         normalising constant is kept so log-densities match the reference's.
GaussianTargetDensity    the same in any dimension <= 8.
evaluate_log() goes through the device (yg_logpost) when a sampler evaluates it; the host
formula below is only used by set-up scripts (meshes, plots).
"""
import numpy as np

from .interface import DensityInterface


class GaussianTargetDensity(DensityInterface):

    def __init__(self, mean, cov, normalised=True):
        self._mean = np.atleast_1d(np.asarray(mean.coefficient if hasattr(mean, 'coefficient') else mean,
                                              dtype=np.float64)).reshape(-1)
        d = self._mean.size
        self._cov = np.asarray(cov, dtype=np.float64).reshape(d, d)
        P = np.linalg.inv(self._cov)
        self._prec = 0.5 * (P + P.T)
        self._logconst = (-0.5 * (d * np.log(2.0 * np.pi) + np.linalg.slogdet(self._cov)[1])) if normalised else 0.0

    @property
    def dimension(self):
        return self._mean.size

    def evaluate_log(self, parameter):
        x = np.asarray(parameter.coefficient, dtype=np.float64).reshape(-1) - self._mean
        return -0.5 * float(x @ (self._prec @ x)) + self._logconst

    def evaluate_on_mesh(self, mesh):
        x = np.asarray(mesh, dtype=np.float64) - self._mean
        return np.exp(-0.5 * np.einsum('...i,ij,...j->...', x, self._prec, x) + self._logconst)

    def device_target(self):
        return dict(g_mean=self._mean, g_prec=self._prec, g_logconst=self._logconst)


class GaussianTargetDensity1d(GaussianTargetDensity):

    def __init__(self, mean, var):
        super().__init__(mean, np.array([[float(var)]]), normalised=False)


class GaussianTargetDensity2d(GaussianTargetDensity):

    def __init__(self, mean, cov):
        super().__init__(mean, cov, normalised=True)
