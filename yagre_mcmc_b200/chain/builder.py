"""ChainBuilder (reference: yagremcmc/chain/builder.py:7-83): property setters + validation, then
build_method() lowers the configured objects to one device problem.  Batched additions:
nChains, seed, device, thin, storeTrajectory, launch (kernel geometry overrides)."""
from abc import ABC, abstractmethod

from .diagnostics import AcceptanceRateDiagnostics


class ChainBuilder(ABC):

    def __init__(self):
        self._bayesModel = None
        self._explicitTarget = None
        self._diagnostics = AcceptanceRateDiagnostics()
        self._nChains = 1
        self._seed = 0
        self._device = None
        self._thin = 1
        self._storeTrajectory = True
        self._launch = None
        self._stateEquality = None

    # ---- reference properties ---------------------------------------------------------------
    @property
    def bayesModel(self):
        return self._bayesModel

    @bayesModel.setter
    def bayesModel(self, model):
        self._bayesModel = model

    @property
    def explicitTarget(self):
        return self._explicitTarget

    @explicitTarget.setter
    def explicitTarget(self, density):
        self._explicitTarget = density

    @property
    def diagnostics(self):
        return self._diagnostics

    @diagnostics.setter
    def diagnostics(self, diagnostics):
        self._diagnostics = diagnostics

    # ---- batched-chain additions ----------------------------------------------------------------
    @property
    def nChains(self):
        return self._nChains

    @nChains.setter
    def nChains(self, n):
        if int(n) < 1:
            raise ValueError("nChains must be a positive integer")
        self._nChains = int(n)

    @property
    def seed(self):
        return self._seed

    @seed.setter
    def seed(self, s):
        self._seed = int(s)

    @property
    def device(self):
        return self._device

    @device.setter
    def device(self, dev):
        self._device = dev

    @property
    def thin(self):
        return self._thin

    @thin.setter
    def thin(self, t):
        if int(t) < 1:
            raise ValueError("thin must be >= 1")
        self._thin = int(t)

    @property
    def storeTrajectory(self):
        return self._storeTrajectory

    @storeTrajectory.setter
    def storeTrajectory(self, flag):
        self._storeTrajectory = bool(flag)

    @property
    def launch(self):
        return self._launch

    @launch.setter
    def launch(self, kw):
        self._launch = dict(kw) if kw else None

    @property
    def stateEquality(self):
        """'exact' (ParameterVector) or 'isclose' (ScalarParameter); None = by dimension of the
        first run's state type is unknown at build time, so 1-d explicit targets default to the
        ScalarParameter rule the reference's 1-d example uses."""
        return self._stateEquality

    @stateEquality.setter
    def stateEquality(self, eq):
        if eq not in (None, 'exact', 'isclose'):
            raise ValueError("stateEquality must be 'exact' or 'isclose'")
        self._stateEquality = eq

    def _common(self):
        return dict(nChains=self._nChains, seed=self._seed, device=self._device, thin=self._thin,
                    storeTrajectory=self._storeTrajectory, launch=self._launch)

    # ---- validation / dispatch ----------------------------------------------------------------
    def validate_target_measure(self):
        if self._bayesModel is None and self._explicitTarget is None:
            raise ValueError("Either bayesian model or explicit target density must be provided for chain setup")
        if self._bayesModel is not None and self._explicitTarget is not None:
            raise ValueError("Only one of bayes model or explicit target density should be provided.")

    def target_is_posterior(self):
        return self._bayesModel is not None

    def target_is_explicit(self):
        return self._explicitTarget is not None

    @abstractmethod
    def _validate_parameters(self):
        ...

    @abstractmethod
    def build_from_model(self):
        ...

    @abstractmethod
    def build_from_target(self):
        ...

    def build_method(self):
        self._validate_parameters()
        self.validate_target_measure()
        if self.target_is_posterior():
            return self.build_from_model()
        if self.target_is_explicit():
            return self.build_from_target()
        raise ValueError("Invalid target distribution")
