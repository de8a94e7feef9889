"""Progress reporting (reference: yagremcmc/utility/verbosity.py:11-52, boilerplate.py:4-27): the
run is cut into slices of max(chainLength // 20, 10) steps; after each slice the rolling
acceptance rate comes from the device counters."""
import logging


def create_logger(name):
    logger = logging.getLogger(name)
    if not logger.handlers:
        h = logging.StreamHandler()
        h.setFormatter(logging.Formatter('%(levelname)s - %(name)s: %(message)s'))
        logger.addHandler(h)
    logger.setLevel(logging.INFO)
    logger.propagate = False
    return logger


class VerbosityController:

    def __init__(self, nPrintIntervals=20, minInterval=10):
        self._n, self._min = nPrintIntervals, minInterval
        self._logger = create_logger(f"MH_{id(self)}")
        self.on = True

    def print_interval(self, chainLength):
        return max(chainLength // self._n, self._min)

    def report(self, stepsDone, diagnostics):
        if not self.on:
            return
        self._logger.info(f"\n\n{stepsDone} steps computed. Calculating diagnostics.\n")
        diagnostics.print_diagnostics(self._logger)
