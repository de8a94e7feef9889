"""Metropolised random walk (reference: yagremcmc/chain/method/mrw.py:9-91): symmetric Gaussian
proposal, acceptance min(1, exp(pi(p) - pi(s))) -- evaluated per chain in the step kernels."""
from ..metropolisHastings import MetropolisHastings
from ..proposal import MRWProposal
from ..target import UnnormalisedPosterior
from ..builder import ChainBuilder
from ..lowering import lower_problem


class MetropolisedRandomWalk(MetropolisHastings):

    def __init__(self, targetDensity, proposalCov, diagnostics, nChains=1, seed=0, device=None, thin=1,
                 storeTrajectory=True, launch=None, equality=None):
        if equality is None:
            equality = 'isclose' if proposalCov.dimension == 1 else 'exact'
        lowered = lower_problem([targetDensity], proposalCov, equality=equality)
        super().__init__(targetDensity, MRWProposal(proposalCov), diagnostics, lowered, nChains=nChains,
                         seed=seed, device=device, thin=thin, storeTrajectory=storeTrajectory, launch=launch)
        if isinstance(targetDensity, UnnormalisedPosterior):
            targetDensity.bind(self._ensemble, 0)


class MRWBuilder(ChainBuilder):

    def __init__(self):
        super().__init__()
        self._proposalCov = None

    @property
    def proposalCovariance(self):
        return self._proposalCov

    @proposalCovariance.setter
    def proposalCovariance(self, covariance):
        self._proposalCov = covariance

    def build_from_model(self):
        target = UnnormalisedPosterior(self._bayesModel.likelihood, self._bayesModel.prior)
        return MetropolisedRandomWalk(target, self._proposalCov, self._diagnostics,
                                      equality=self._stateEquality or 'exact', **self._common())

    def build_from_target(self):
        return MetropolisedRandomWalk(self._explicitTarget, self._proposalCov, self._diagnostics,
                                      equality=self._stateEquality, **self._common())

    def _validate_parameters(self):
        if self._proposalCov is None:
            raise ValueError("Proposal Covariance not set for MRW")
