"""CPU: host logic of the drop-in API -- builder validation (reference test/test_builder.py),
lowering of reference-style objects to plain arrays, covariance operators, sharding maths."""
import numpy as np
import pytest

import bench_problems as bp
from golden_io import load
from yagre_mcmc_b200.parameter import ParameterVector, ScalarParameter
from yagre_mcmc_b200.model import ForwardModel, LinearModelSolver, LotkaVolterraRK4Solver, LotkaVolterraParameter
from yagre_mcmc_b200.model.interface import SolverInterface
from yagre_mcmc_b200.statistics import (IIDCovarianceMatrix, DiagonalCovarianceMatrix, DenseCovarianceMatrix,
                                        Gaussian, CentredGaussianNoise, Data, AdditiveGaussianNoiseLikelihood,
                                        BayesianRegressionModel, BayesianRegressionModelHierarchy,
                                        GaussianTargetDensity2d, GaussianTargetDensity1d)
from yagre_mcmc_b200.statistics.interface import DensityInterface
from yagre_mcmc_b200.utility import Hierarchy, SharedComponent
from yagre_mcmc_b200.chain.method import MRWBuilder, MLDABuilder, AMBuilder, PCNBuilder, AEMBuilder
from yagre_mcmc_b200.statistics import AEMLikelihood, AEMNoise
from yagre_mcmc_b200.chain.target import UnnormalisedPosterior
from yagre_mcmc_b200.chain.lowering import lower_problem
from yagre_mcmc_b200.parallel import shard_range, moments_from_stats


def test_builder_validation_messages():
    b = MRWBuilder()
    with pytest.raises(ValueError, match="Proposal Covariance not set for MRW"):      # test_builder.py:20
        b.build_method()
    b.proposalCovariance = IIDCovarianceMatrix(2, 1.0)
    with pytest.raises(ValueError, match="Either bayesian model or explicit target"):
        b.build_method()
    b.explicitTarget = GaussianTargetDensity2d(ParameterVector(np.zeros(2)), np.eye(2))
    b.bayesModel = object()
    with pytest.raises(ValueError, match="Only one of bayes model or explicit target"):
        b.build_method()
    m = MLDABuilder()
    with pytest.raises(ValueError, match="Coarse proposal covariance not set for MLDA"):
        m.build_method()
    m.baseProposalCovariance = IIDCovarianceMatrix(2, 1.0)
    with pytest.raises(ValueError, match="Subchain lengths not set for MLDA"):
        m.build_method()
    m.subChainLengths = [3]
    m.explicitTarget = GaussianTargetDensity2d(ParameterVector(np.zeros(2)), np.eye(2))
    with pytest.raises(ValueError, match="Surrogate targets not set for MLDA"):
        m.build_method()
    m.surrogateTargets = [m.explicitTarget, m.explicitTarget]
    with pytest.raises(ValueError, match="Number of sub-chain lengths does not match number of surrogate"):
        m.build_method()
    a = AMBuilder()
    with pytest.raises(ValueError, match="Regularisation parameter must be non-negative"):
        a.regularisationParameter = -1.0


def test_pcn_builder_validation():
    """reference chain/method/pcn.py:42-46,80-88."""
    b = PCNBuilder()
    with pytest.raises(ValueError, match="Step size not set in PCN"):
        b.build_method()
    b.stepSize = 0.01
    b.explicitTarget = GaussianTargetDensity2d(ParameterVector(np.zeros(2)), np.eye(2))
    with pytest.raises(RuntimeError, match="only defined in relation to a Bayesian model"):
        b.build_method()
    lik = AdditiveGaussianNoiseLikelihood(Data(np.zeros((2, 2))), ForwardModel(LinearModelSolver(np.eye(2), np.zeros(2))),
                                          CentredGaussianNoise(IIDCovarianceMatrix(2, 1.)))
    b2 = PCNBuilder()
    b2.stepSize = 0.01
    b2.bayesModel = BayesianRegressionModel(lik, Gaussian(ParameterVector(np.array([0.1, 0.])), IIDCovarianceMatrix(2, 1.)))
    with pytest.raises(ValueError, match="requires centred prior"):
        b2.build_method()
    b2.stepSize = 0.7
    with pytest.raises(AssertionError):
        b2.build_method()


def test_aem_builder_validation():
    """reference chain/method/aem.py:66-79, statistics/likelihood.py:99-100, statistics/noise.py:29-33."""
    data = Data(np.zeros((3, 2)))
    fm = ForwardModel(LinearModelSolver(np.eye(2), np.zeros(2)))
    noise = CentredGaussianNoise(IIDCovarianceMatrix(2, 0.3))
    with pytest.raises(ValueError, match="Smallest senisible data size for AEM is 2"):
        AEMLikelihood(data, fm, noise, 1)
    with pytest.raises(NotImplementedError, match="independent measurement noise"):
        AEMLikelihood(data, fm, CentredGaussianNoise(DenseCovarianceMatrix(np.eye(2) + 0.1)), 5)
    prior = Gaussian(ParameterVector(np.zeros(2)), IIDCovarianceMatrix(2, 1.))
    plain = AdditiveGaussianNoiseLikelihood(data, fm, noise)
    aem = AEMLikelihood(data, fm, noise, 10, True)
    assert aem.device_aem() == dict(min_data=10, heuristic=True) and isinstance(aem.noiseModel, AEMNoise)
    assert AEMNoise.scaling_heuristic(np.array([0.5, 0.01])) == 100 and AEMNoise.scaling_heuristic(np.array([0.2, 0.1])) == 4.0
    b = AEMBuilder()
    b.baseProposalCovariance = IIDCovarianceMatrix(2, 0.5)
    b.subChainLengths = [5]
    b.explicitTarget = GaussianTargetDensity2d(ParameterVector(np.zeros(2)), np.eye(2))
    b.surrogateTargets = [b.explicitTarget]
    with pytest.raises(NotImplementedError, match="only makes sense if the target emerges from a Bayesian model"):
        b.build_method()
    b2 = AEMBuilder()
    b2.baseProposalCovariance = IIDCovarianceMatrix(2, 0.5)
    b2.subChainLengths = [5]
    b2.bayesModel = BayesianRegressionModelHierarchy(Hierarchy([aem, plain]), SharedComponent(prior, 2))
    with pytest.raises(ValueError, match="Likelihood on level 1 is not adaptive"):
        b2.build_method()


def lv_objects():
    meta, arr = bp.lv_problem(True)
    design, data = arr["L0_design"], arr["L0_data"]
    cfg = dict(T=10., alpha=0.8, gamma=0.4, nData=10, dataDim=2)
    noise = CentredGaussianNoise(IIDCovarianceMatrix(2, 0.04))
    prior = Gaussian(LotkaVolterraParameter.from_coefficient(np.zeros(2)), IIDCovarianceMatrix(2, 1.4))
    lik = [AdditiveGaussianNoiseLikelihood(Data(data), ForwardModel(LotkaVolterraRK4Solver(design, dict(cfg, rk4Steps=N))), noise)
           for N in (64, 512)]
    return meta, arr, lik, prior


def test_lowering_of_lv_hierarchy_matches_plain_arrays():
    meta, arr, lik, prior = lv_objects()
    hier = BayesianRegressionModelHierarchy(Hierarchy(lik), SharedComponent(prior, 2))
    posts = [UnnormalisedPosterior(hier.level(k).likelihood, hier.level(k).prior) for k in range(2)]
    low = lower_problem(posts, IIDCovarianceMatrix(2, 0.1), subChainLength=3)
    assert (low.model, low.dim, low.levels, low.J) == ('lv', 2, 2, 3)
    for k, v in arr.items():
        assert np.array_equal(low.arrays[k], v), k         # bit-identical lowering


def test_lowering_matches_golden_fixture_arrays():
    """The reference-side lowering used for the fixtures and the product lowering agree bitwise."""
    meta, a = load("mlda_linear")
    noise = CentredGaussianNoise(IIDCovarianceMatrix(2, np.sqrt(0.3) ** 2))
    prior = Gaussian(ParameterVector(a["L0_prior_mean"]), IIDCovarianceMatrix(2, 5.0))
    lik = [AdditiveGaussianNoiseLikelihood(Data(a["L0_data"]), ForwardModel(LinearModelSolver(a[f"L{l}_G"], a[f"L{l}_b"])), noise)
           for l in range(2)]
    posts = [UnnormalisedPosterior(l, prior) for l in lik]
    low = lower_problem(posts, IIDCovarianceMatrix(2, 0.5), subChainLength=5)
    for k in ("prop_L", "L0_noise_prec", "L1_prior_prec", "L0_G", "L1_b", "L0_data"):
        assert np.array_equal(low.arrays[k], a[k]), k
    meta, a = load("mrw_gauss2d_dense")
    low = lower_problem([GaussianTargetDensity2d(ParameterVector(a["L0_g_mean"]), bp.GAUSS2D_COV)],
                        DenseCovarianceMatrix([[1.2, -0.3], [-0.3, 0.5]]))
    np.testing.assert_allclose(low.arrays["prop_L"], a["prop_L"], rtol=1e-15)
    np.testing.assert_allclose(low.arrays["L0_g_prec"], a["L0_g_prec"], rtol=1e-13)
    np.testing.assert_allclose(low.arrays["L0_g_logconst"], a["L0_g_logconst"], rtol=1e-14)


def test_unrecognised_plugins_are_refused_without_fallback():
    class MySolver(SolverInterface):
        status = evaluation = None
        def interpolate(self, p): pass
        def invoke(self): pass

    class MyDensity(DensityInterface):
        def evaluate_log(self, p): return 0.0
    lik = AdditiveGaussianNoiseLikelihood(Data(np.zeros((2, 2))), ForwardModel(MySolver()),
                                          CentredGaussianNoise(IIDCovarianceMatrix(2, 1.)))
    prior = Gaussian(ParameterVector(np.zeros(2)), IIDCovarianceMatrix(2, 1.))
    with pytest.raises(NotImplementedError, match="no device implementation"):
        lower_problem([UnnormalisedPosterior(lik, prior)], IIDCovarianceMatrix(2, 1.))
    with pytest.raises(NotImplementedError, match="no device implementation"):
        lower_problem([MyDensity()], IIDCovarianceMatrix(2, 1.))
    with pytest.raises(NotImplementedError, match="levels"):
        g = GaussianTargetDensity2d(ParameterVector(np.zeros(2)), np.eye(2))
        lower_problem([g, g, g, g], IIDCovarianceMatrix(2, 1.))          # three surrogates crash in the reference too
    g = GaussianTargetDensity2d(ParameterVector(np.zeros(2)), np.eye(2))
    low = lower_problem([g, g, g], IIDCovarianceMatrix(2, 1.), subChainLength=4)      # two surrogates: mlda.py:112-117
    assert low.levels == 3 and low.J == 4 and "L2_g_mean" in low.arrays
    with pytest.raises(ValueError, match="centred Gaussian noise"):
        AdditiveGaussianNoiseLikelihood(Data(np.zeros((2, 2))), None, object())


def test_covariance_operators_match_reference_fixture():
    _, a = load("postprocessing")
    dc = DenseCovarianceMatrix(a["dense_C"])
    for x, y1, y2, n2 in zip(a["dense_v"], a["dense_chol_apply"], a["dense_inv_apply"], a["dense_norm2"]):
        np.testing.assert_allclose(dc.apply_chol_factor(x), y1, rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(dc.apply_inverse(x), y2, rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(dc.induced_norm_squared(x), n2, rtol=1e-12)
        np.testing.assert_allclose(x @ dc.precision() @ x, n2, rtol=1e-12)
    d = DiagonalCovarianceMatrix(np.array([0.1, 0.3]))
    assert np.array_equal(d.precision(), np.diag(np.reciprocal(np.array([0.1, 0.3]))))
    assert d.precision()[0, 1] == 0.0
    assert IIDCovarianceMatrix(3, 0.5).dimension == 3


def test_parameter_types():
    p = ParameterVector(np.array([1., 2.]))
    assert p == ParameterVector(np.array([1., 2.])) and not (p == ParameterVector(np.array([1., 2. + 1e-15])))
    with pytest.raises(ValueError):
        p.clone_with([1., 2.])
    s = ScalarParameter(np.array([1.0]))
    assert s == ScalarParameter(np.array([1.0 + 1e-12]))         # math.isclose (scalar.py:38-43)
    assert s.equality == 'isclose' and p.equality == 'exact'
    with pytest.raises(Exception):
        ScalarParameter(1.0)
    lv = LotkaVolterraParameter.from_interpolation(np.array([0.4, 0.6]))
    np.testing.assert_allclose(lv.evaluate(), [0.4, 0.6])
    stacked = ParameterVector(np.zeros((7, 2)))
    assert stacked.nChains == 7 and stacked.dimension == 2


def test_hierarchy_containers():
    h = Hierarchy(['c', 'm', 'f'])
    assert h.level(-1) == 'f' and h.level(0) == 'c' and h.size == 3
    with pytest.raises(ValueError):
        h.level(3)
    s = SharedComponent('x', 2)
    assert s.level(1) == 'x'
    with pytest.raises(ValueError, match="mismatched sizes"):
        BayesianRegressionModelHierarchy(Hierarchy([1, 2]), SharedComponent(0, 3))
    with pytest.raises(RuntimeError):
        BayesianRegressionModel(Hierarchy([1, 2]), 0)


def test_shard_ranges_partition_the_ensemble():
    for n, g in [(524288, 8), (65536, 3), (10, 4), (7, 7)]:
        r = [shard_range(n, k, g) for k in range(g)]
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(r[k][1] == r[k + 1][0] for k in range(g - 1))
        assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_pooled_moments_and_rhat_from_sufficient_statistics():
    rng = np.random.default_rng(5)
    C, n, d = 64, 500, 2
    x = rng.standard_normal((C, n, d)) * np.array([1.0, 3.0]) + np.array([2.0, -1.0]) + 0.2 * rng.standard_normal((C, 1, d))
    mean_c = x.mean(axis=1)
    m2_c = np.einsum('cni,cnj->cij', x - mean_c[:, None], x - mean_c[:, None])
    var_c = x.var(axis=1, ddof=1)
    vec = np.concatenate([[C, n, 1234.0], mean_c.sum(0), np.einsum('ci,cj->ij', mean_c, mean_c).ravel(),
                          m2_c.sum(0).ravel(), var_c.sum(0)])
    out = moments_from_stats(vec, d)
    flat = x.reshape(-1, d)
    np.testing.assert_allclose(out["mean"], flat.mean(0), rtol=1e-12)
    np.testing.assert_allclose(out["covariance"], np.cov(flat.T), rtol=1e-10)
    W = var_c.mean(0)
    B_over_n = mean_c.var(axis=0, ddof=1)
    np.testing.assert_allclose(out["rhat"], np.sqrt(((n - 1) / n * W + B_over_n) / W), rtol=1e-12)
    assert abs(out["acceptance_rate"] - 1234.0 / (C * n)) < 1e-15


def test_trajectory_file_format_roundtrip(tmp_path):
    """io.py on the host: [length, d, n_chains] float64 .npy + JSON side-car."""
    from yagre_mcmc_b200 import io
    x = np.random.default_rng(0).standard_normal((7, 2, 5))
    f = io.save_trajectory(str(tmp_path / "t.npy"), x, thin=3, chain_offset=40)
    y, side = io.load_trajectory(f)
    assert np.array_equal(x, y) and side["thin"] == 3 and side["layout"] == "[length, d, n_chains]"
    assert io.as_reference_layout(y).shape == (7, 5, 2)
    with pytest.raises(ValueError):
        io.save_trajectory(str(tmp_path / "bad"), np.zeros((3, 3), dtype=np.float32))


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under yagre_mcmc_b200/ may import, load or execute it
    (no CPU fallback), and the package must not read /root/reference."""
    import os
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "yagre_mcmc_b200")
    bad = []
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b|libyagre_oracle|/root/reference|yagremcmc\b(?!/)", txt, re.M) and f.endswith(".py"):
                    if re.search(r"^\s*(from|import)\s+(oracle|yagremcmc)\b|libyagre_oracle", txt, re.M):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad
