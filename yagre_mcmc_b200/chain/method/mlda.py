"""Multi-level delayed acceptance (reference: yagremcmc/chain/method/mlda.py:12-344).

One surrogate (`nSurrogates == 1` branch :102-110, MLDA._acceptance_probability :146-154): per fine step
the device runs subChainLengths[0] coarse MRW steps from the current state, skips the fine model when the
sub-chain did not move (metropolisHastings.py:60-61), and otherwise screens the sub-chain's end point with
    min(1, exp(pi_f(p) + pi_c(s) - pi_c(p) - pi_f(s))).
Two surrogates, exactly what the reference does (:12-43,60-71,112-117; pinned by tests/golden/mlda3_*.npz):
MLDAProposal hands the state to the top SurrogateTransition, whose generate_proposal() runs the BASE MRW
for subChainLengths[1] steps (subChainLengths[0] is never read) and returns its end point; the screen is
MLDA._acceptance_probability with the FINEST surrogate as pi_c.  Three and five surrogates crash in the
reference (AttributeError: MetropolisedRandomWalk has no set_state), four are not implemented here.

Extension: `baseProposalCovariance` may be an AdaptiveCovarianceMatrix descriptor (chain/adaptive.py); the
coarse MRW then adapts per chain on the device -- the reference's AdaptiveMRWProposal as the proposal method
of the surrogate MRW, update() before every coarse proposal (pinned by tests/golden/am_mlda_*.npz)."""
from ..metropolisHastings import MetropolisHastings
from ..proposal import ProposalMethod
from ..target import UnnormalisedPosterior
from ..builder import ChainBuilder
from ..diagnostics import AcceptanceRateDiagnostics, DummyDiagnostics
from ..lowering import lower_problem
from ..adaptive import AdaptiveCovarianceMatrix
from ...utility.hierarchy import Hierarchy


class MLDAProposal(ProposalMethod):
    """Descriptor of the composite proposal: a coarse MRW sub-chain."""

    def __init__(self, surrogateTargets, baseProposalCov, nSteps):
        super().__init__()
        self._surrogates = list(surrogateTargets)
        self._cov = baseProposalCov
        self._nSteps = list(nSteps)

    @property
    def nSurrogates(self):
        return len(self._surrogates)

    def surrogate(self, sIdx):
        if sIdx < -1 or sIdx >= len(self._surrogates):
            raise IndexError(f"invalid surrogate index: {sIdx}")
        return self._surrogates[sIdx]

    @property
    def covariance(self):
        return self._cov

    @property
    def subChainLengths(self):
        return self._nSteps


class MLDA(MetropolisHastings):

    def __init__(self, targetDensity, surrogateDensities, baseProposalCov, nSteps, targetDiagnostics,
                 surrogateDiagnosticsList, nChains=1, seed=0, device=None, thin=1, storeTrajectory=True,
                 launch=None, equality='exact', aem=None):
        nSur = len(surrogateDensities)
        if nSur not in (1, 2):
            raise NotImplementedError(
                f"{nSur} surrogates: the device runs MLDA with one or two surrogates (the reference crashes for "
                "three and five, mlda.py:23-33)")
        # mlda.py:100-117: one surrogate -> subChainLengths[0] MRW steps; two -> the top SurrogateTransition
        # runs the base MRW for ITS length subChainLengths[1]
        J = nSteps[0] if nSur == 1 else nSteps[1]
        adaptive, cov = None, baseProposalCov
        if isinstance(baseProposalCov, AdaptiveCovarianceMatrix):
            adaptive, cov = baseProposalCov.device_config(), baseProposalCov.covariance
        levels = list(surrogateDensities) + [targetDensity]
        lowered = lower_problem(levels, cov, subChainLength=J, equality=equality)
        proposal = MLDAProposal(surrogateDensities, baseProposalCov, nSteps)
        super().__init__(targetDensity, proposal, targetDiagnostics, lowered, nChains=nChains, seed=seed,
                         device=device, thin=thin, storeTrajectory=storeTrajectory, launch=launch, aem=aem,
                         adaptive=adaptive)
        self._surrogateDiagnostics = surrogateDiagnosticsList
        self._J = J
        for lvl, t in enumerate(levels):
            if isinstance(t, UnnormalisedPosterior):
                t.bind(self._ensemble, lvl)

    @property
    def nSurrogates(self):
        return self._proposalMethod.nSurrogates

    def surrogate(self, sIdx):
        return self._proposalMethod.surrogate(sIdx)

    def _update_diagnostics(self):
        """+ the base surrogate's diagnostics (reference: the MRW of SurrogateHierarchy level 0 records every coarse
        step, mlda.py:58-62): the kernels count the accepted coarse sub-steps of the ensemble; per-chain coarse rates
        are not kept.  With two surrogates the second surrogate's diagnostics see nothing, as in the reference (its
        _accept_reject is never called, mlda.py:112-117)."""
        super()._update_diagnostics()
        dg = (self._surrogateDiagnostics or [None])[0]
        if dg is not None and hasattr(dg, "process_totals"):
            c = self._ensemble.counters()
            dg.process_totals(c["coarse_accepted"], c["transitions"] * self._J)

    def evaluation_counts(self):
        """Forward evaluations actually performed (coarse, fine) -- skipped ones do not count."""
        c = self._ensemble.counters()
        return c['coarse_evals'], c['fine_evals']

    def proposal_factors(self):
        """Current per-chain coarse proposal factors L [nChains_local, d, d] (adaptive base covariance only)."""
        return self._ensemble.state()['prop_L'].permute(2, 0, 1).cpu().numpy()


class MLDABuilder(ChainBuilder):

    def __init__(self):
        super().__init__()
        self._basePropCov = None
        self._nSteps = None
        self._surrTgts = None
        self._surrDgnstList = None
        self._tgtDgnst = None
        self._biasCorrection = None

    @property
    def baseProposalCovariance(self):
        return self._basePropCov

    @baseProposalCovariance.setter
    def baseProposalCovariance(self, cov):
        self._basePropCov = cov

    @property
    def subChainLengths(self):
        return self._nSteps

    @subChainLengths.setter
    def subChainLengths(self, nSteps):
        self._nSteps = nSteps

    @property
    def surrogateTargets(self):
        return self._surrTgts

    @surrogateTargets.setter
    def surrogateTargets(self, tgts):
        self._surrTgts = tgts

    @property
    def targetDiagnostics(self):
        return self._tgtDgnst

    @targetDiagnostics.setter
    def targetDiagnostics(self, diagnostics):
        self._tgtDgnst = diagnostics

    @property
    def surrogateDiagnostics(self):
        return self._surrDgnstList

    @surrogateDiagnostics.setter
    def surrogateDiagnostics(self, lst):
        self._surrDgnstList = lst

    @property
    def biasCorrection(self):
        return self._biasCorrection

    @biasCorrection.setter
    def biasCorrection(self, bias):
        self._biasCorrection = bias

    def _validate_parameters(self):
        if self._basePropCov is None:
            raise ValueError("Coarse proposal covariance not set for MLDA")
        if self._nSteps is None:
            raise ValueError("Subchain lengths not set for MLDA")
        if self._biasCorrection is not None:
            raise NotImplementedError("BiasCorrection is broken in the reference (chain/target.py:59-67) "
                                      "and has no device implementation")
        if self._bayesModel is not None:
            if not isinstance(self._bayesModel, Hierarchy):
                raise ValueError("MLDA requires a hierarchy of models.")
            if self._surrTgts is not None:
                raise ValueError("Cannot set explicit surrogate targets for a hierarchy of Bayesian models.")
            if len(self._nSteps) != self._bayesModel.size - 1:
                raise ValueError("Number of sub-chain lengths does not match the size of the model hierarchy.")
            if self._surrDgnstList is not None and len(self._surrDgnstList) != self._bayesModel.size - 1:
                raise ValueError("Number of diagnostics does not match the size of the model hierarchy")
        if self._explicitTarget is not None:
            if self._surrTgts is None:
                raise ValueError("Surrogate targets not set for MLDA")
            if len(self._nSteps) != len(self._surrTgts):
                raise ValueError("Number of sub-chain lengths does not match number of surrogate targets")
            if self._surrDgnstList is not None and len(self._surrDgnstList) != len(self._surrTgts):
                raise ValueError("Number of diagnostics does not match the size of the model hierarchy")

    def create_diagnostics(self, nSurrogates):
        if self._tgtDgnst is None:
            self._tgtDgnst = AcceptanceRateDiagnostics()
        if self._surrDgnstList is None:
            self._surrDgnstList = [DummyDiagnostics() for _ in range(nSurrogates)]

    def build_from_model(self):
        self._validate_parameters()
        n = self._bayesModel.size
        posts = [UnnormalisedPosterior(self._bayesModel.level(k).likelihood, self._bayesModel.level(k).prior)
                 for k in range(n)]
        self.create_diagnostics(n - 1)
        return self.build_mlda(posts[-1], posts[:-1], self._basePropCov, self._nSteps, self._tgtDgnst,
                               self._surrDgnstList, equality=self._stateEquality or 'exact', **self._common())

    def build_from_target(self):
        self.create_diagnostics(len(self._surrTgts))
        return self.build_mlda(self._explicitTarget, list(self._surrTgts), self._basePropCov, self._nSteps,
                               self._tgtDgnst, self._surrDgnstList, equality=self._stateEquality or 'exact',
                               **self._common())

    def build_mlda(self, tgtPost, surPost, bpc, nS, tgtD, surD, **kw):      # reference mlda.py:337-338
        return MLDA(tgtPost, surPost, bpc, nS, tgtD, surD, **kw)
