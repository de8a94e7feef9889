"""Adaptive proposal covariance (reference interface: yagremcmc/chain/adaptive.py:8-64).

The reference ships only the interface -- `update()` is called in set_state(), i.e. before
every proposal, with the chain's newest state (:55-60) -- and no working implementation
(chain/method/deprecated/am.py:152 raises).  The device implements the Haario-type rule
described in DESIGN.md ("Adaptive Metropolis") per chain inside generic_kernel.cu; this class
is its descriptor."""
from ..statistics.interface import CovarianceOperatorInterface


class AdaptiveCovarianceMatrix(CovarianceOperatorInterface):

    def __init__(self, initCov, idleSteps, collectionSteps, eps, scale=None, refresh=1):
        if initCov.dimension == 1:
            raise NotImplementedError("Adaptivity not implemented for scalar chains.")   # adaptive.py:41-43
        self._cov = initCov
        self.idleSteps = int(idleSteps)
        self.collectionSteps = int(collectionSteps)
        self.eps = float(eps)
        self.scale = scale
        self.refresh = int(refresh)

    @property
    def dimension(self):
        return self._cov.dimension

    @property
    def covariance(self):
        return self._cov

    def apply_chol_factor(self, x):
        return self._cov.apply_chol_factor(x)

    def apply_inverse(self, x):
        return self._cov.apply_inverse(x)

    def update(self):
        raise NotImplementedError("the update runs per chain on the device, before every proposal")

    def device_config(self):
        return dict(idle=self.idleSteps, collection=self.collectionSteps, eps=self.eps,
                    scale=self.scale or 0.0, refresh=self.refresh)
