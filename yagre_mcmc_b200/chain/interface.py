"""Chain diagnostics ABC (reference: yagremcmc/chain/interface.py:4-16).

In the reference `process(transitionData)` is called once per step of the single chain.  The
batched sampler keeps the counters on the device and hands a whole run to
`process_ensemble(stats)` instead; `process` remains for API compatibility."""
from abc import ABC, abstractmethod


class ChainDiagnostics(ABC):

    def process(self, transitionData):
        raise NotImplementedError("batched chains report through process_ensemble()")

    @abstractmethod
    def process_ensemble(self, stats):
        ...

    @abstractmethod
    def print_diagnostics(self, logger):
        ...

    @abstractmethod
    def reset(self):
        ...
