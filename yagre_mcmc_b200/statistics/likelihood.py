"""Additive Gaussian-noise likelihood (reference: yagremcmc/statistics/likelihood.py:13-87):
logL = -1/2 sum_rows || F(theta) - d_row ||^2_{Sigma^-1}.

The reference memoises the last five evaluations (:51,61-72); forward models are
deterministic, so the kernels simply carry log pi(state) in the chain state instead.
"""
from .interface import DensityInterface
from .noise import CentredGaussianNoise, AEMNoise


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


class AEMLikelihood(AdditiveGaussianNoiseLikelihood):
    """Likelihood with an adaptive error model (reference: yagremcmc/statistics/likelihood.py:90-155).
    Descriptor of what aem_mh_kernel does per chain: Welford of F_fine - F_coarse on accepted fine
    steps; from `minDataSize` errors on the residual is shifted by their mean (:140-145), from
    minDataSize + 1 on the noise variance is inflated by their variance (:147-155, noise.py:41-54)."""

    def __init__(self, data, forwardModel, noiseModel, minDataSize, useNoiseHeuristic=False):
        if minDataSize < 2:
            raise ValueError("Smallest senisible data size for AEM is 2.")
        self._minDataSize = int(minDataSize)
        super().__init__(data, forwardModel, AEMNoise(noiseModel, useNoiseHeuristic))

    @property
    def minDataSize(self):
        return self._minDataSize

    @property
    def useNoiseHeuristic(self):
        return self._noiseModel.useHeuristic

    def device_aem(self):
        return dict(min_data=self._minDataSize, heuristic=self._noiseModel.useHeuristic)
