"""Diagnostics (reference: yagremcmc/chain/diagnostics.py:8-107).

AcceptanceRateDiagnostics  global acceptance rate = accepted / transitions (:44-46), pooled
                           over the ensemble; per-chain rates via acceptance_rates().
FullDiagnostics            + Welford mean / marginal variance of the PRE-transition states
                           (:91-94), accumulated inside the step kernels.
"""
import numpy as np

from .interface import ChainDiagnostics
from ..statistics.estimation import WelfordAccumulator


class DummyDiagnostics(ChainDiagnostics):

    def process_ensemble(self, stats):
        return

    def process_totals(self, n_accept_total, n_decisions_total):
        return

    def print_diagnostics(self, logger):
        return

    def reset(self):
        return


class AcceptanceRateDiagnostics(ChainDiagnostics):

    def __init__(self):
        self._lag = None
        self.reset()

    @property
    def lag(self):
        return self._lag

    @lag.setter
    def lag(self, value):
        if value <= 0:
            raise ValueError("Lag must be a positive integer.")
        self._lag = value

    def process_ensemble(self, stats):
        """stats: n_accept[nChains] (cumulative), transitions per chain (cumulative)."""
        prev_acc, prev_n = self._n_accept, self._n_transitions
        self._n_accept = np.asarray(stats['n_accept'], dtype=np.int64)
        self._n_transitions = int(stats['transitions'])
        if prev_acc is not None and self._n_transitions > prev_n:
            self._rolling = float((self._n_accept - prev_acc).sum()) / (
                (self._n_transitions - prev_n) * self._n_accept.size)
        elif self._n_transitions:
            self._rolling = self.global_acceptance_rate()

    def process_totals(self, n_accept_total, n_decisions_total):
        """Ensemble totals only (the coarse sub-chains of delayed acceptance live inside the step kernels, which
        count their accepted sub-steps but keep no per-chain record): feeds global_acceptance_rate()."""
        prev = self._totals
        self._totals = (int(n_accept_total), int(n_decisions_total))
        if prev is not None and self._totals[1] > prev[1]:
            self._rolling = (self._totals[0] - prev[0]) / float(self._totals[1] - prev[1])
        elif self._totals[1]:
            self._rolling = self._totals[0] / float(self._totals[1])

    def acceptance_rates(self):
        if self._n_accept is None or not self._n_transitions:
            return None
        return self._n_accept / float(self._n_transitions)

    def global_acceptance_rate(self):
        if self._totals is not None:
            return self._totals[0] / float(self._totals[1]) if self._totals[1] else 0.0
        if self._n_accept is None or not self._n_transitions:
            return 0.0
        return float(self._n_accept.sum()) / (self._n_transitions * self._n_accept.size)

    def rolling_acceptance_rate(self):
        if self._rolling is None:
            raise RuntimeError("Insufficient data for rolling acceptance rate.")
        return self._rolling

    def print_diagnostics(self, logger):
        try:
            logger.info(f"  - Rolling acceptance rate: {self.rolling_acceptance_rate():.4f}")
        except RuntimeError as e:
            logger.warning(f"  - Rolling acceptance rate unavailable: {e}")

    def reset(self):
        self._n_accept = None
        self._n_transitions = 0
        self._rolling = None
        self._totals = None


class FullDiagnostics(ChainDiagnostics):

    def __init__(self):
        self._diagnostics = AcceptanceRateDiagnostics()
        self._accumulator = WelfordAccumulator()
        self._lag = None
        self._squeeze = False

    @property
    def lag(self):
        return self._lag

    @lag.setter
    def lag(self, lag):
        self._lag = lag
        self._diagnostics.lag = lag

    def process_ensemble(self, stats):
        self._diagnostics.process_ensemble(stats)
        self._squeeze = stats.get('squeeze', False)
        self._accumulator.load(stats['welford_n'], stats['w_mean'], stats['w_m2_diag'])

    def process_totals(self, n_accept_total, n_decisions_total):
        self._diagnostics.process_totals(n_accept_total, n_decisions_total)

    def global_acceptance_rate(self):
        return self._diagnostics.global_acceptance_rate()

    def acceptance_rates(self):
        return self._diagnostics.acceptance_rates()

    def rolling_acceptance_rate(self):
        return self._diagnostics.rolling_acceptance_rate()

    def mean(self):
        m = self._accumulator.mean()
        return m[0] if self._squeeze else m

    def marginal_variance(self):
        v = self._accumulator.marginal_variance()
        return v[0] if self._squeeze else v

    def print_diagnostics(self, logger):
        self._diagnostics.print_diagnostics(logger)
        try:
            cn = np.asarray(self._accumulator.condition_number())
            logger.info(f"  - Estimated condition number: {float(np.median(cn)):.4f}")
        except RuntimeError as e:
            logger.warning(f"  - Condition number unavailable: {e}")

    def reset(self):
        self._diagnostics.reset()
        self._accumulator.reset()
