"""Unnormalised posterior (reference: yagremcmc/chain/target.py:4-22): log-likelihood + log-prior
density.  Tempered / bias-corrected variants (:25-67) are out of scope (broken in the
reference).

evaluate_log() is served by the device (yg_logpost) once a sampler has bound the target."""
import numpy as np

from ..statistics.interface import DensityInterface


class UnnormalisedPosterior(DensityInterface):

    def __init__(self, likelihood, prior):
        self._likelihood = likelihood
        self._prior = prior
        self._binding = None           # (ensemble, level) set by the sampler

    @property
    def likelihood(self):
        return self._likelihood

    @property
    def prior(self):
        return self._prior

    def bind(self, ensemble, level):
        self._binding = (ensemble, level)

    def evaluate_log(self, parameter):
        if self._binding is None:
            raise RuntimeError("posterior not bound to a device ensemble yet: build the chain first "
                               "(there is no CPU evaluation path)")
        ens, level = self._binding
        coef = np.asarray(parameter.coefficient if hasattr(parameter, 'coefficient') else parameter,
                          dtype=np.float64)
        out = ens.logpost(level, np.atleast_2d(coef)).cpu().numpy()
        return float(out[0]) if coef.ndim == 1 else out

    def device_level(self):
        model, lvl = self._likelihood.device_level()
        lvl['prior_mean'] = np.asarray(self._prior.mean.coefficient, dtype=np.float64).reshape(-1)
        lvl['prior_prec'] = self._prior.covariance.precision()
        return model, lvl
