"""Additive Gaussian-noise likelihood (reference: yagremcmc/statistics/likelihood.py:13-87):
logL = -1/2 sum_rows || F(theta) - d_row ||^2_{Sigma^-1}.

The reference memoises the last five evaluations (:51,61-72); forward models are
deterministic, so the kernels simply carry log pi(state) in the chain state instead.
"""
from .interface import DensityInterface
from .noise import CentredGaussianNoise


class AdditiveGaussianNoiseLikelihood(DensityInterface):

    def __init__(self, data, forwardModel, noiseModel):
        if not isinstance(noiseModel, CentredGaussianNoise):
            raise ValueError("AdditiveGaussianNoiseLikelihood requires centred Gaussian noise.")
        self._data = data
        self._fwdModel = forwardModel
        self._noiseModel = noiseModel

    @property
    def data(self):
        return self._data

    @property
    def forwardModel(self):
        return self._fwdModel

    @property
    def noiseModel(self):
        return self._noiseModel

    def evaluate_log(self, parameter):
        raise NotImplementedError(
            "likelihoods are evaluated on the device as part of the posterior: use "
            "UnnormalisedPosterior(likelihood, prior).evaluate_log(parameter)")

    def device_level(self):
        """Plain arrays of this likelihood (one hierarchy level)."""
        solver = self._fwdModel.solver
        model = solver.device_model()              # raises NotImplementedError for unknown plugins
        lvl = dict(solver.device_level())
        lvl['data'] = self._data.array
        lvl['noise_prec'] = self._noiseModel.covariance.precision()
        return model, lvl
