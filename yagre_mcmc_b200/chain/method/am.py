"""Adaptive Metropolis (builder surface of the reference's dead implementation,
yagremcmc/chain/method/deprecated/am.py:162-230: idleSteps, collectionSteps,
regularisationParameter, initialCovariance; interface chain/adaptive.py:37-64).

The recurrence is ours (DESIGN.md "Adaptive Metropolis", parity unpinned vs the reference,
pinned GPU <-> oracle): per chain, full-matrix Welford of the states seen from step
`idleSteps` on, fed BEFORE each proposal; once `collectionSteps` states are collected the
proposal factor becomes chol(s (Cov + eps I)), s = 2.4^2 / d, refreshed every `refresh` steps."""
from ..metropolisHastings import MetropolisHastings
from ..proposal import MRWProposal
from ..target import UnnormalisedPosterior
from ..builder import ChainBuilder
from ..adaptive import AdaptiveCovarianceMatrix
from ..lowering import lower_problem


class AdaptiveMRWProposal(MRWProposal):

    def __init__(self, adaptiveCov):
        super().__init__(adaptiveCov)


class AdaptiveMetropolis(MetropolisHastings):

    def __init__(self, targetDensity, initCov, idleSteps, collectionSteps, regParam, diagnostics,
                 scale=None, refresh=1, nChains=1, seed=0, device=None, thin=1, storeTrajectory=True, launch=None):
        acov = AdaptiveCovarianceMatrix(initCov, idleSteps, collectionSteps, regParam, scale=scale, refresh=refresh)
        lowered = lower_problem([targetDensity], initCov, equality='exact')
        super().__init__(targetDensity, AdaptiveMRWProposal(acov), diagnostics, lowered, nChains=nChains,
                         seed=seed, device=device, adaptive=acov.device_config(), thin=thin,
                         storeTrajectory=storeTrajectory, launch=launch)
        if isinstance(targetDensity, UnnormalisedPosterior):
            targetDensity.bind(self._ensemble, 0)

    def proposal_factors(self):
        """Current per-chain proposal factors L [nChains_local, d, d] (device -> host)."""
        return self._ensemble.state()['prop_L'].permute(2, 0, 1).cpu().numpy()


class AMBuilder(ChainBuilder):

    def __init__(self):
        super().__init__()
        self._idleSteps = None
        self._collectionSteps = None
        self._regularisationParameter = None
        self._initialCovariance = None
        self._scale = None
        self._refresh = 1

    @property
    def idleSteps(self):
        return self._idleSteps

    @idleSteps.setter
    def idleSteps(self, iSteps):
        self._idleSteps = iSteps

    @property
    def collectionSteps(self):
        return self._collectionSteps

    @collectionSteps.setter
    def collectionSteps(self, cSteps):
        self._collectionSteps = cSteps

    @property
    def regularisationParameter(self):
        return self._regularisationParameter

    @regularisationParameter.setter
    def regularisationParameter(self, eps):
        if eps < 0:
            raise ValueError("Regularisation parameter must be non-negative.")
        self._regularisationParameter = eps

    @property
    def initialCovariance(self):
        return self._initialCovariance

    @initialCovariance.setter
    def initialCovariance(self, cov):
        self._initialCovariance = cov

    @property
    def scaling(self):
        return self._scale

    @scaling.setter
    def scaling(self, s):
        self._scale = s

    @property
    def refreshInterval(self):
        return self._refresh

    @refreshInterval.setter
    def refreshInterval(self, r):
        self._refresh = int(r)

    def _validate_parameters(self):
        for name, v in (("idleSteps", self._idleSteps), ("collectionSteps", self._collectionSteps),
                        ("regularisationParameter", self._regularisationParameter),
                        ("initialCovariance", self._initialCovariance)):
            if v is None:
                raise ValueError(f"{name} not set for adaptive Metropolis")

    def _build(self, target):
        return AdaptiveMetropolis(target, self._initialCovariance, self._idleSteps, self._collectionSteps,
                                  self._regularisationParameter, self._diagnostics, scale=self._scale,
                                  refresh=self._refresh, **self._common())

    def build_from_model(self):
        return self._build(UnnormalisedPosterior(self._bayesModel.likelihood, self._bayesModel.prior))

    def build_from_target(self):
        return self._build(self._explicitTarget)
