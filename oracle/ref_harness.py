"""
TEST INFRASTRUCTURE, CONTAINER-ONLY.  Drives the UNMODIFIED reference
(/root/reference, rkutri/yagre-mcmc) under injected noise to produce golden
trajectories (see make_golden.py).  Nothing here is imported by the product,
by `-m gpu` tests, by smoke() or by bench.py: /root/reference does not exist on
the GPU box.

What is ours and what is the reference's:
  * the chain stack (MRWBuilder / MLDABuilder -> MetropolisHastings.run,
    yagremcmc/chain/metropolisHastings.py:103-120, chain/method/mlda.py:100-154,
    statistics/likelihood.py:33-84, statistics/gaussian.py:19-66) is the
    reference's, untouched;
  * RK4LotkaVolterraSolver is OUR SolverInterface plugin
    (yagremcmc/model/interface.py:7-67): same design/initial-condition semantics
    and flow association order as yagremcmc/test/testSetup.py:96-141, but a
    fixed-step classical RK4 instead of scipy.solve_ivp, as north_star
    prescribes.  The reference has no RK4 code, so the RK4 arithmetic order
    defined here IS the specification the C oracle and the CUDA kernels follow;
  * NoiseInjector replaces the two module-level RNG entry points
    (yagremcmc.statistics.gaussian.standard_normal, gaussian.py:2,63 and
    yagremcmc.chain.metropolisHastings.uniform, metropolisHastings.py:2,68) so
    the host decides z and u; arrays are indexed by (fine step, sub step), not
    by call count, so a skipped uniform does not shift the stream.
"""
import os
import sys
import contextlib
import io

import numpy as np

REFERENCE_ROOT = os.environ.get("YAGRE_REFERENCE_ROOT", "/root/reference")


def import_reference():
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True          # /root/reference is read-only
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import yagremcmc  # noqa: F401
    return yagremcmc


import_reference()

from yagremcmc.model.interface import SolverInterface            # noqa: E402
from yagremcmc.model.evaluation import EvaluationStatus          # noqa: E402
from yagremcmc.model.forwardModel import ForwardModel            # noqa: E402
from yagremcmc.parameter.vector import ParameterVector           # noqa: E402
from yagremcmc.parameter.scalar import ScalarParameter           # noqa: E402
from yagremcmc.statistics.data import Data                       # noqa: E402
from yagremcmc.statistics.covariance import (                    # noqa: E402
    IIDCovarianceMatrix, DiagonalCovarianceMatrix, DenseCovarianceMatrix)
from yagremcmc.statistics.gaussian import Gaussian               # noqa: E402
from yagremcmc.statistics.noise import CentredGaussianNoise      # noqa: E402
from yagremcmc.statistics.likelihood import AdditiveGaussianNoiseLikelihood  # noqa: E402
from yagremcmc.statistics.bayesModel import BayesianRegressionModel          # noqa: E402
from yagremcmc.statistics.modelHierarchy import BayesianRegressionModelHierarchy  # noqa: E402
from yagremcmc.utility.hierarchy import SharedComponent, Hierarchy           # noqa: E402
from yagremcmc.chain.method.mrw import MRWBuilder, MetropolisedRandomWalk   # noqa: E402
from yagremcmc.chain.method.mlda import MLDABuilder              # noqa: E402
from yagremcmc.chain.method.pcn import PCNBuilder                # noqa: E402
from yagremcmc.chain.method.aem import AEMBuilder                # noqa: E402
from yagremcmc.statistics.likelihood import AEMLikelihood        # noqa: E402
from yagremcmc.chain.diagnostics import FullDiagnostics          # noqa: E402
from yagremcmc.test.testSetup import (                           # noqa: E402
    LotkaVolterraParameter, GaussianTargetDensity1d, GaussianTargetDensity2d)
import yagremcmc.statistics.gaussian as _ref_gaussian            # noqa: E402
import yagremcmc.chain.metropolisHastings as _ref_mh             # noqa: E402


# --------------------------------------------------------------------------
# forward-model plugins
# --------------------------------------------------------------------------

class RK4LotkaVolterraSolver(SolverInterface):
    """Fixed-step classical RK4 for x' = a x - b x y, y' = d x y - g y.

    One evaluation = nData independent integrations over [0, T] (one per design
    row = initial condition) sharing (b, d) = exp(theta); output row n = state
    at T (testSetup.py:118-141).  Arithmetic order (numpy, unfused):
        fx = alpha*x - beta*x*y   == (alpha*x) - ((beta*x)*y)   testSetup.py:98
        fy = delta*x*y - gamma*y  == ((delta*x)*y) - (gamma*y)  testSetup.py:99
        h = T/N ; h2 = 0.5*h ; h6 = h/6.0
        k1 = f(s); k2 = f(s + h2*k1); k3 = f(s + h2*k2); k4 = f(s + h*k3)
        s <- s + h6*(((k1 + 2.0*k2) + 2.0*k3) + k4)
    Non-finite outputs are mapped element-wise to +inf (=> logL = -inf =>
    rejected by the unchanged acceptance rule; SURVEY section 7 "hard parts").
    """

    def __init__(self, design, config):
        self.x_ = np.asarray(design, dtype=np.float64)
        self.T_ = float(config['T'])
        self.N_ = int(config['rk4Steps'])
        self.fixedParam_ = [float(config['alpha']), float(config['gamma'])]
        self.dataShape_ = (config['nData'], config['dataDim'])
        self.param_ = [None, None]
        self.evaluation_ = None
        self.status_ = EvaluationStatus.NONE
        self.nEvaluations = 0

    @property
    def status(self):
        return self.status_

    @property
    def evaluation(self):
        return self.evaluation_

    @property
    def dataShape(self):
        return self.dataShape_

    def interpolate(self, parameter):
        paramEval = parameter.evaluate()        # exp(coefficient), testSetup.py:57-58
        self.param_ = [paramEval[0], paramEval[1]]

    def invoke(self):
        alpha, gamma = self.fixedParam_
        beta, delta = self.param_
        x = self.x_[:, 0].copy()
        y = self.x_[:, 1].copy()
        h = self.T_ / self.N_
        h2 = 0.5 * h
        h6 = h / 6.0

        def f(x, y):
            return alpha * x - beta * x * y, delta * x * y - gamma * y

        with np.errstate(all='ignore'):
            for _ in range(self.N_):
                k1x, k1y = f(x, y)
                k2x, k2y = f(x + h2 * k1x, y + h2 * k1y)
                k3x, k3y = f(x + h2 * k2x, y + h2 * k2y)
                k4x, k4y = f(x + h * k3x, y + h * k3y)
                x = x + h6 * (((k1x + 2.0 * k2x) + 2.0 * k3x) + k4x)
                y = y + h6 * (((k1y + 2.0 * k2y) + 2.0 * k3y) + k4y)

        ev = np.stack([x, y], axis=1)
        ev = np.where(np.isfinite(ev), ev, np.inf)
        self.evaluation_ = ev
        self.status_ = EvaluationStatus.SUCCESS
        self.nEvaluations += 1


class LinearSolver(SolverInterface):
    """G.theta + b, arithmetic of exampleSetup.py:46 (`A @ theta + b`).

    Restated here (not imported) because exampleSetup.py sits outside the
    yagremcmc package next to scripts that import matplotlib.
    """

    def __init__(self, A, b):
        self._A = np.asarray(A, dtype=np.float64)
        self._b = np.asarray(b, dtype=np.float64)
        self._coef = None
        self._evaluation = None
        self._status = EvaluationStatus.NONE

    @property
    def status(self):
        return self._status

    @property
    def evaluation(self):
        return self._evaluation

    def interpolate(self, parameter):
        self._coef = parameter.coefficient

    def invoke(self):
        self._evaluation = self._A @ self._coef + self._b
        self._status = EvaluationStatus.SUCCESS


# --------------------------------------------------------------------------
# noise injection
# --------------------------------------------------------------------------

class NoiseInjector:
    """z[nSteps, J, d], u_c[nSteps, J], u_f[nSteps] (single level: J = 1 and
    only u_f is read).  `level` is set by wrappers around the reference's
    _accept_reject so the uniform can be attributed to the coarse sub-chain or
    to the fine screen even when an equality skip swallowed a coarse uniform.
    """

    def __init__(self, z, u_c, u_f):
        self.z = np.asarray(z, dtype=np.float64)
        self.u_c = None if u_c is None else np.asarray(u_c, dtype=np.float64)
        self.u_f = np.asarray(u_f, dtype=np.float64)
        self.J = self.z.shape[1]
        self.kz = 0
        self.cur = (0, 0)
        self.level = 'fine'
        self.log = []       # RNG call order, e.g. 'NUNUNUU'

    def standard_normal(self, size=None):
        n, j = divmod(self.kz, self.J)
        self.kz += 1
        self.cur = (n, j)
        self.log.append('N')
        out = self.z[n, j].copy()
        assert out.size == size
        return out

    def uniform(self, low=0., high=1., size=1):
        assert low == 0. and high == 1. and size == 1
        n, j = self.cur
        if self.level == 'coarse':
            self.log.append('u')
            return np.array([self.u_c[n, j]])
        self.log.append('U')
        return np.array([self.u_f[n]])

    @contextlib.contextmanager
    def installed(self):
        saved = (_ref_gaussian.standard_normal, _ref_mh.uniform)
        _ref_gaussian.standard_normal = self.standard_normal
        _ref_mh.uniform = self.uniform
        try:
            yield self
        finally:
            _ref_gaussian.standard_normal, _ref_mh.uniform = saved


def _tag_level(method, injector, level):
    """Wrap method._accept_reject (instance attribute) to tag the RNG level."""
    inner = method._accept_reject

    def wrapped(proposal, state):
        prev = injector.level
        injector.level = level
        try:
            return inner(proposal, state)
        finally:
            injector.level = prev

    method._accept_reject = wrapped


def quiet(fn, *a, **kw):
    """The reference prints banners from constructors (mrw.py:45, mlda.py:126)."""
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **kw)


# --------------------------------------------------------------------------
# running the reference chain under injection
# --------------------------------------------------------------------------

def run_reference_chain(sampler, initState, nTransitions, injector, twoLevel):
    """Returns trajectory [nTransitions+1, d], accepted [nTransitions] (0/1)."""
    if twoLevel:
        _tag_level(sampler, injector, 'fine')
        _tag_level(sampler.surrogate(0), injector, 'coarse')
    else:
        _tag_level(sampler, injector, 'fine')
    with injector.installed():
        sampler.run(nTransitions + 1, initState, verbose=False)
    traj = np.array([np.asarray(s, dtype=np.float64).reshape(-1)
                     for s in sampler.chain.trajectory])
    dg = sampler.diagnostics
    decisions = dg._diagnostics._decisions if isinstance(dg, FullDiagnostics) \
        else dg._decisions
    return traj, np.asarray(decisions, dtype=np.uint8)


def covariance_from_spec(kind, value, dim):
    if kind == 'iid':
        return IIDCovarianceMatrix(dim, float(value))
    if kind == 'diag':
        return DiagonalCovarianceMatrix(np.asarray(value, dtype=np.float64))
    if kind == 'dense':
        return DenseCovarianceMatrix(np.asarray(value, dtype=np.float64))
    raise ValueError(kind)


# --------------------------------------------------------------------------
# adaptive Metropolis through the reference's LIVE interface
# --------------------------------------------------------------------------
from yagremcmc.chain.adaptive import AdaptiveCovarianceMatrix, AdaptiveMRWProposal   # noqa: E402
from yagremcmc.chain.metropolisHastings import MetropolisHastings                   # noqa: E402


class HaarioAdaptiveCovariance(AdaptiveCovarianceMatrix):
    """OUR concrete AdaptiveCovarianceMatrix (the reference ships only the abstract class,
    chain/adaptive.py:8-34, and a dead implementation, chain/method/deprecated/am.py).

    What the reference pins: update() takes no argument, sees the chain through set_chain()
    (adaptive.py:20-21) and is called by AdaptiveMRWProposal.set_state (adaptive.py:55-60), i.e.
    by MetropolisHastings.run BEFORE every proposal (metropolisHastings.py:117), at which point
    chain.trajectory[-1] is the current state; the covariance it returns is swapped into the MRW
    proposal for the proposal drawn right after.

    What is ours: with t the number of update() calls so far, for
    t >= idle:  n = t - idle + 1;  delta = x - mean;  mean += delta / n;
                M2 += outer(delta, x - mean)                    (Welford, full matrix)
    and when n >= collection, n >= 2 and (n - collection) % refresh == 0:
                C = scale * (0.5 (M2 + M2') / (n - 1) + eps I),  scale = 2.4^2 / d by default;
    the proposal covariance becomes DenseCovarianceMatrix(C) (scipy Cholesky, covariance.py:69-86)
    unless C is not positive definite, in which case the previous one is kept.
    """

    def __init__(self, initCov, idleSteps, collectionSteps, eps, scale=None, refresh=1):
        super().__init__(initCov)
        d = initCov.dimension
        self.idle, self.collection, self.refresh = int(idleSteps), int(collectionSteps), int(refresh)
        self.eps = float(eps)
        self.scale = 2.4 * 2.4 / d if not scale else float(scale)
        self.t = 0
        self.mean = np.zeros(d)
        self.M2 = np.zeros((d, d))
        self.nRefresh = 0

    def update(self):
        x = np.asarray(self._chain.trajectory[-1], dtype=np.float64).reshape(-1)
        t = self.t
        self.t += 1
        if t < self.idle:
            return
        n = t - self.idle + 1
        delta = x - self.mean
        self.mean = self.mean + delta / n
        self.M2 = self.M2 + np.outer(delta, x - self.mean)
        if n >= self.collection and n >= 2 and (n - self.collection) % self.refresh == 0:
            cov = 0.5 * (self.M2 + self.M2.T) / (n - 1)
            C = self.scale * (cov + self.eps * np.eye(x.size))
            try:
                self._cov = DenseCovarianceMatrix(C)
                self.nRefresh += 1
            except np.linalg.LinAlgError:
                pass

    def chol_factor(self):
        """Current lower factor L (p = s + L z), in the lowered form of make_golden.lower_proposal."""
        c = self._cov
        if isinstance(c, DenseCovarianceMatrix):
            return np.array(c.cholFactor_)
        return np.diag(np.sqrt(np.reciprocal(c._precision)))


class AdaptiveMetropolis(MetropolisHastings):
    """Single-level adaptive MRW assembled from the reference's own parts: the unmodified
    MetropolisHastings.run loop, the unmodified AdaptiveMRWProposal and the MRW acceptance rule
    (chain/method/mrw.py:51-57)."""

    def __init__(self, targetDensity, adaptiveCov, diagnostics):
        super().__init__(targetDensity, AdaptiveMRWProposal(adaptiveCov), diagnostics)
        adaptiveCov.set_chain(self._chain)

    _acceptance_probability = MetropolisedRandomWalk._acceptance_probability


def make_surrogate_adaptive(mlda, adaptiveCov):
    """Two-level delayed acceptance with an adaptive coarse proposal: the MRW surrogate of an
    unmodified MLDA object (mlda.py:58-62) gets the reference's AdaptiveMRWProposal as its proposal
    method, bound to the surrogate's own chain.  update() is then called before every COARSE
    proposal with the sub-chain's current state (metropolisHastings.py:117 of the surrogate's run)."""
    sur = mlda.surrogate(0)
    sur._proposalMethod = AdaptiveMRWProposal(adaptiveCov)
    adaptiveCov.set_chain(sur.chain)
    return mlda
