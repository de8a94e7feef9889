"""Preconditioned Crank-Nicolson (reference: yagremcmc/chain/method/pcn.py:9-88).

Proposal  sqrt(1 - 2h) * state + sqrt(2h) * xi,  xi ~ prior  (:23-35); the acceptance ratio is the
LIKELIHOOD ratio (:52-57, the target density handed to MetropolisHastings is the likelihood);
the prior must be a centred Gaussian (:44-46).  On the device this is the MRW step kernel with
the pCN proposal rule (yg_problem.proposal = YG_PROPOSAL_PCN) and a zero prior precision."""
import numpy as np

from ..metropolisHastings import MetropolisHastings
from ..proposal import ProposalMethod
from ..target import UnnormalisedPosterior
from ..builder import ChainBuilder
from ..lowering import lower_problem
from ...statistics.gaussian import Gaussian


class PCNProposal(ProposalMethod):

    def __init__(self, prior, stepSize):
        if not isinstance(prior, Gaussian):
            raise NotImplementedError("PCN only supports Gaussian priors")
        super().__init__()
        self.prior_ = prior
        self._stepSize = stepSize

    @property
    def stepSize(self):
        return self._stepSize

    @property
    def prior(self):
        return self.prior_


class LikelihoodTarget(UnnormalisedPosterior):
    """The pCN target: log-likelihood only.  Lowered as a regression level with zero prior
    precision, so the kernels' `logL + logprior` adds an exact -0.0."""

    def __init__(self, likelihood, prior):
        super().__init__(likelihood, prior)

    def device_level(self):
        model, lvl = super().device_level()
        lvl['prior_prec'] = np.zeros_like(np.asarray(lvl['prior_prec'], dtype=np.float64))
        return model, lvl


class PreconditionedCrankNicolson(MetropolisHastings):

    def __init__(self, likelihood, prior, stepSize, diagnostics, nChains=1, seed=0, device=None, thin=1,
                 storeTrajectory=True, launch=None, equality='exact'):
        assert 0 < stepSize and stepSize <= 0.5                           # pcn.py:42
        proposalMethod = PCNProposal(prior, stepSize)
        mean = np.asarray(prior.mean.coefficient, dtype=np.float64).reshape(-1)
        if np.any(mean != 0.0):
            raise ValueError("Preconditioned Crank Nicholson requires centred prior")   # pcn.py:44-46
        target = LikelihoodTarget(likelihood, prior)
        lowered = lower_problem([target], prior.covariance, equality=equality, proposal='pcn',
                                pcnStep=float(stepSize), pcnMean=mean)
        super().__init__(target, proposalMethod, diagnostics, lowered, nChains=nChains, seed=seed,
                         device=device, thin=thin, storeTrajectory=storeTrajectory, launch=launch)
        target.bind(self._ensemble, 0)


class PCNBuilder(ChainBuilder):

    def __init__(self):
        super().__init__()
        self._stepSize = None

    @property
    def stepSize(self):
        return self._stepSize

    @stepSize.setter
    def stepSize(self, h):
        self._stepSize = h

    def build_from_model(self):
        return PreconditionedCrankNicolson(self._bayesModel.likelihood, self._bayesModel.prior, self._stepSize,
                                           self._diagnostics, equality=self._stateEquality or 'exact',
                                           **self._common())

    def build_from_target(self):
        raise RuntimeError("PCN is only defined in relation to a Bayesian model")

    def _validate_parameters(self):
        if self._stepSize is None:
            raise ValueError("Step size not set in PCN.")
