"""Welford accumulator (reference: yagremcmc/statistics/estimation.py:4-58).

In the batched backend the recurrence runs inside the step kernels (per chain, full second
moment matrix); this class is the host-side VIEW of those accumulators with the reference's
accessor names, so FullDiagnostics keeps its interface."""
import numpy as np


class WelfordAccumulator:

    def __init__(self):
        self.reset()

    def load(self, nData, mean, m2_diag):
        """mean / m2_diag: [nChains, d] arrays copied from the device."""
        self._dataSize = int(nData)
        self._mean = np.asarray(mean)
        self._welfordM2 = np.asarray(m2_diag)

    @property
    def nData(self):
        return self._dataSize

    def mean(self):
        return self._mean

    def marginal_variance(self):
        if self._dataSize < 2:
            raise RuntimeError("Insufficient data for variance estimation.")
        return self._welfordM2 / (self._dataSize - 1)

    def condition_number(self):
        mv = self.marginal_variance()
        if np.min(mv) < 1e-12:
            raise RuntimeError("Singular marginal variance.")
        return np.max(mv, axis=-1) / np.min(mv, axis=-1)

    def reset(self):
        self._dataSize = 0
        self._mean = None
        self._welfordM2 = None
