"""Adaptive error model MLDA (reference: yagremcmc/chain/method/aem.py:7-82).

Two-level delayed acceptance in which, on every ACCEPTED fine step, the difference between the fine
and the coarse forward evaluation of the accepted proposal feeds the coarse likelihood's error model
(:44-56).  The device kernel (aem_mh_kernel, generic_kernel.cu) keeps one error model per chain, as a
batch of independent reference chains would, and reproduces the coarse likelihood's LRU(3) cache
(utility/memoisation.py:76-149) including the fact that cached values survive model updates."""
import numpy as np

from .mlda import MLDA, MLDABuilder
from ...statistics.likelihood import AEMLikelihood


class AdaptiveErrorModel(MLDA):

    def __init__(self, targetDensity, surrogateDensities, baseProposalCov, nSteps, targetDiagnostics,
                 surrogateDiagnostics, **kw):
        coarse = surrogateDensities[0].likelihood
        super().__init__(targetDensity, surrogateDensities, baseProposalCov, nSteps, targetDiagnostics,
                         surrogateDiagnostics, aem=coarse.device_aem(), **kw)

    def error_model(self):
        """Per-chain state of the coarse likelihood's error model: nData [n], mean [n, dataDim],
        marginal variance [n, dataDim] (NaN while fewer than two errors were seen) -- the
        WelfordAccumulator behind AEMLikelihood.accumulator in the reference."""
        st = self._ensemble.state()
        n = st['aem_n'].cpu().numpy()
        mean = st['aem_mean'].t().cpu().numpy()
        with np.errstate(all='ignore'):
            var = np.where(n[:, None] > 1, st['aem_m2'].t().cpu().numpy() / (n[:, None] - 1.0), np.nan)
        return dict(nData=n, mean=mean, marginal_variance=var)


class AEMBuilder(MLDABuilder):

    def _validate_parameters(self):
        super()._validate_parameters()
        if self._bayesModel is None:
            raise NotImplementedError("Adaptive error correction only makes "
                                      "sense if the target emerges from a Bayesian model.")
        for i in range(self._bayesModel.size):
            if not isinstance(self._bayesModel.level(i).likelihood, AEMLikelihood):
                raise ValueError(f"Likelihood on level {i} is not adaptive.")

    def build_mlda(self, tgtPost, surPost, bpc, nS, tgtD, surD, **kw):
        return AdaptiveErrorModel(tgtPost, surPost, bpc, nS, tgtD, surD, **kw)
