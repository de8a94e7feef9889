"""Proposal descriptors (reference: yagremcmc/chain/proposal.py:4-27, chain/method/mrw.py:9-38).

In the batched backend a proposal method does not generate anything on the host: it
describes what the kernels do (covariance factor, sub-chain structure)."""
from abc import ABC


class ProposalMethod(ABC):

    def __init__(self):
        self._state = None
        self._stateType = None

    def get_state(self):
        return self._state

    def set_state(self, newState):
        self._state = newState
        self._stateType = type(newState)

    @property
    def stateType(self):
        return self._stateType

    def generate_proposal(self):
        raise NotImplementedError("proposals are drawn on the device (Philox4x32-10 + Box-Muller)")


class MRWProposal(ProposalMethod):
    """Gaussian random walk around the state with a fixed covariance."""

    def __init__(self, proposalCov):
        super().__init__()
        self._cov = proposalCov

    @property
    def covariance(self):
        return self._cov

    @covariance.setter
    def covariance(self, cov):
        self._cov = cov
