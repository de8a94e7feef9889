"""Unnormalised posterior (reference: yagremcmc/chain/target.py:4-22): log-likelihood + log-prior
density, and its tempered variant (:25-43): tempering * log-likelihood + log-prior.  BiasCorrection
(:46-67) is not mirrored: it crashes in the reference (it passes a raw array where a parameter object is
expected).

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


class TemperedUnnormalisedPosterior(UnnormalisedPosterior):
    """reference chain/target.py:25-43.  Usable as an explicit (surrogate) target of MLDABuilder; the
    reference's TemperedMLDA wrapper itself (chain/method/tmlda.py:44-52) cannot be constructed."""

    def __init__(self, likelihood, prior, tempering):
        super().__init__(likelihood, prior)
        self.tempering = tempering

    @property
    def tempering(self):
        return self._tempering

    @tempering.setter
    def tempering(self, value):
        if not 0.0 <= float(value) <= 1.0:          # chain/method/tmlda.py:24-29
            raise ValueError(f"Invalid tempering parameter: {value} (must be in [0, 1]).")
        self._tempering = float(value)

    def device_level(self):
        model, lvl = super().device_level()
        lvl['tempering'] = np.array(self._tempering)
        return model, lvl
