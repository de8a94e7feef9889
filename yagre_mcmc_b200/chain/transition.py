"""Outcome codes of one transition (reference: yagremcmc/chain/transition.py:1-25).  The kernels
write exactly these codes into the optional `accepted` output of yg_run."""


class TransitionData:

    REJECTED = 0
    ACCEPTED = 1

    def __init__(self, state, proposal, outcome):
        if outcome not in (TransitionData.REJECTED, TransitionData.ACCEPTED):
            raise RuntimeError(f"invalid MC transition outcome: {outcome}")
        self._state, self._proposal, self._outcome = state, proposal, outcome

    @property
    def state(self):
        return self._state

    @property
    def proposal(self):
        return self._proposal

    @property
    def outcome(self):
        return self._outcome
