"""
TEST INFRASTRUCTURE, CONTAINER-ONLY: generates tests/golden/*.npz by running the
UNMODIFIED reference (via oracle/ref_harness.py) under injected noise.

    python oracle/make_golden.py            # rewrites every fixture (about 2 min)

The reference holds no known-answer trajectories, so these
fixtures ARE the pins for the C oracle (oracle/yagre_oracle.c) and, through it
and directly, for the CUDA path.  Each .npz carries the lowered problem
(exactly the arrays the C-ABI takes), the injected noise, and the reference's
outputs: trajectory, accept decisions, log-posterior per level along the
trajectory, Welford mean / marginal variance where FullDiagnostics was used.

Constants are those of the BASELINE.json configs:
  C1 example_mcmc_1d.py:12-30              C2 example_mcmc_2d_singleLevel.py:19-27
  2-D two level example_mcmc_2d_twoLevel.py:10-40
  C3 example_inference_linearModel_twoLevel.py:33-74,128-129,173
  C4/C5 example_inference_lotkaVolterra_{single,two}Level.py:29-106
"""
import json
import os
import sys

import numpy as np
from numpy.random import Generator, Philox

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_harness as rh   # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


# --------------------------------------------------------------------------
# lowering helpers: reference objects -> plain arrays (what the C-ABI takes)
# --------------------------------------------------------------------------

def lower_proposal(kind, value, dim):
    """Lower-triangular factor L with p = s + L z.  Diagonal kinds mimic
    DiagonalCovarianceMatrix exactly: it stores reciprocal(var) and applies
    sqrt(reciprocal(precision)) (covariance.py:37-38,51-52)."""
    if kind == 'iid':
        var = np.full(dim, float(value))
        return np.diag(np.sqrt(np.reciprocal(np.reciprocal(var))))
    if kind == 'diag':
        var = np.asarray(value, dtype=np.float64)
        return np.diag(np.sqrt(np.reciprocal(np.reciprocal(var))))
    if kind == 'dense':
        from scipy.linalg import cholesky
        return cholesky(np.asarray(value, dtype=np.float64), lower=True)   # covariance.py:78
    raise ValueError(kind)


def diag_precision(var, dim):
    """IID/Diagonal covariance -> dense precision matrix with exact zeros off
    the diagonal (covariance.py:37-38,54-55)."""
    var = np.full(dim, float(var)) if np.ndim(var) == 0 else np.asarray(var, dtype=np.float64)
    return np.diag(np.reciprocal(var))


def make_noise(rng, nChains, nSteps, J, d, zero_at=(), u_edge=True):
    z = rng.standard_normal((nChains, nSteps, J, d))
    u_c = rng.random((nChains, nSteps, J))
    u_f = rng.random((nChains, nSteps))
    for (c, n, j) in zero_at:           # equality-skip edge cases (metropolisHastings.py:60-61)
        if j is None:
            z[c, n, :, :] = 0.0
        else:
            z[c, n, j, :] = 0.0
    if u_edge and nSteps > 12:
        u_f[0, 5] = 0.0                 # u == 0 accepts whenever a >= 0
        u_f[0, 9] = np.nextafter(1.0, 0.0)
        u_c[0, 7, 0] = 0.0
    return z, u_c, u_f


def logpost_along(target, stateType, traj):
    cache = {}
    out = np.empty(len(traj))
    for i, s in enumerate(traj):
        key = s.tobytes()
        if key not in cache:
            cache[key] = float(np.asarray(target.evaluate_log(stateType(s.copy()))).reshape(-1)[0])
        out[i] = cache[key]
    return out


def save(name, meta, arrays):
    os.makedirs(OUT, exist_ok=True)
    arrays = {k: np.asarray(v) for k, v in arrays.items()}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), meta=json.dumps(meta), **arrays)
    print(f"  wrote {name}.npz  ({sum(a.nbytes for a in arrays.values())/1024:.0f} KiB raw)")


# --------------------------------------------------------------------------
# Gaussian explicit targets
# --------------------------------------------------------------------------

def gauss2d_level_arrays(mean, cov):
    from scipy.stats import multivariate_normal
    mean = np.asarray(mean, dtype=np.float64)
    cov = np.asarray(cov, dtype=np.float64)
    prec = np.linalg.inv(cov)
    prec = 0.5 * (prec + prec.T)
    # scipy's logpdf = -0.5*(d log 2pi + logdet) - 0.5 maha   (testSetup.py:36-40)
    logconst = float(multivariate_normal(mean, cov).logpdf(mean))
    return mean, prec, logconst


def case_gauss1d():
    nChains, nSteps = 3, 300
    rng = Generator(Philox(101))
    z, u_c, u_f = make_noise(rng, nChains, nSteps, 1, 1, zero_at=[(1, 20, 0)])
    tgt = rh.GaussianTargetDensity1d(rh.ScalarParameter(np.array([1.5])), 1.)
    propVar = 1.5
    traj, acc, lp, wm, wv = [], [], [], [], []
    for c in range(nChains):
        inj = rh.NoiseInjector(z[c], None, u_f[c])
        mcmc = rh.quiet(rh.MetropolisedRandomWalk, tgt, rh.IIDCovarianceMatrix(1, propVar), rh.FullDiagnostics())
        t, a = rh.run_reference_chain(mcmc, rh.ScalarParameter(np.array([-3.])), nSteps, inj, False)
        traj.append(t); acc.append(a)
        lp.append(logpost_along(tgt, rh.ScalarParameter, t))
        wm.append(np.asarray(mcmc.diagnostics.mean()).reshape(-1))
        wv.append(np.asarray(mcmc.diagnostics.marginal_variance()).reshape(-1))
    meta = dict(model='gauss', dim=1, levels=1, J=1, eq='isclose',
                note='C1 example_mcmc_1d.py:12-30; ScalarParameter equality is math.isclose (scalar.py:38-43)')
    save("mrw_gauss1d", meta, dict(
        prop_L=lower_proposal('iid', propVar, 1),
        L0_g_mean=[1.5], L0_g_prec=[[1.0]], L0_g_logconst=0.0,
        theta0=np.full((nChains, 1), -3.0), z=z, u_c=u_c, u_f=u_f,
        traj=traj, accepted=acc, logpost_L0=lp, welford_mean=wm, welford_var=wv))


TGT_MEAN = np.array([1., 1.5])
TGT_COV = np.array([[2.4, -0.5], [-0.5, 0.7]])


def case_gauss2d(name, propKind, propValue, seed):
    nChains, nSteps = 3, 300
    rng = Generator(Philox(seed))
    z, u_c, u_f = make_noise(rng, nChains, nSteps, 1, 2, zero_at=[(2, 11, 0)])
    tgt = rh.GaussianTargetDensity2d(rh.ParameterVector(TGT_MEAN), TGT_COV)
    traj, acc, lp, wm, wv = [], [], [], [], []
    for c in range(nChains):
        inj = rh.NoiseInjector(z[c], None, u_f[c])
        b = rh.MRWBuilder()
        b.explicitTarget = tgt
        b.proposalCovariance = rh.covariance_from_spec(propKind, propValue, 2)
        b.diagnostics = rh.FullDiagnostics()
        mcmc = rh.quiet(b.build_method)
        t, a = rh.run_reference_chain(mcmc, rh.ParameterVector(np.array([-8., -7.])), nSteps, inj, False)
        traj.append(t); acc.append(a)
        lp.append(logpost_along(tgt, rh.ParameterVector, t))
        wm.append(mcmc.diagnostics.mean()); wv.append(mcmc.diagnostics.marginal_variance())
    m, P, lc = gauss2d_level_arrays(TGT_MEAN, TGT_COV)
    meta = dict(model='gauss', dim=2, levels=1, J=1, eq='exact',
                note=f'C2 example_mcmc_2d_singleLevel.py:19-27, proposal {propKind}')
    save(name, meta, dict(
        prop_L=lower_proposal(propKind, propValue, 2),
        L0_g_mean=m, L0_g_prec=P, L0_g_logconst=lc,
        theta0=np.tile([-8., -7.], (nChains, 1)), z=z, u_c=u_c, u_f=u_f,
        traj=traj, accepted=acc, logpost_L0=lp, welford_mean=wm, welford_var=wv))


def case_mlda_gauss2d():
    nChains, nSteps, J = 3, 250, 6
    rng = Generator(Philox(303))
    z, u_c, u_f = make_noise(rng, nChains, nSteps, J, 2,
                             zero_at=[(0, 3, None), (1, 14, 2), (2, 30, None)])
    surMean = TGT_MEAN + np.array([0.15, -0.2])
    surCov = 2. * TGT_COV + np.array([[0.1, 1.2], [1.2, 0.05]])
    tgt = rh.GaussianTargetDensity2d(rh.ParameterVector(TGT_MEAN), TGT_COV)
    sur = rh.GaussianTargetDensity2d(rh.ParameterVector(surMean), surCov)
    traj, acc, lpc, lpf, order = [], [], [], [], []
    for c in range(nChains):
        inj = rh.NoiseInjector(z[c], u_c[c], u_f[c])
        b = rh.MLDABuilder()
        b.explicitTarget = tgt
        b.surrogateTargets = [sur]
        b.baseProposalCovariance = rh.IIDCovarianceMatrix(2, 1.)
        b.subChainLengths = [J]
        mcmc = rh.quiet(b.build_method)
        t, a = rh.run_reference_chain(mcmc, rh.ParameterVector(np.array([-8., -7.])), nSteps, inj, True)
        traj.append(t); acc.append(a)
        lpf.append(logpost_along(tgt, rh.ParameterVector, t))
        lpc.append(logpost_along(sur, rh.ParameterVector, t))
        order.append(''.join(inj.log))
    m1, P1, c1 = gauss2d_level_arrays(TGT_MEAN, TGT_COV)
    m0, P0, c0 = gauss2d_level_arrays(surMean, surCov)
    meta = dict(model='gauss', dim=2, levels=2, J=J, eq='exact', rng_order=order,
                note='example_mcmc_2d_twoLevel.py:10-40')
    save("mlda_gauss2d", meta, dict(
        prop_L=lower_proposal('iid', 1., 2),
        L0_g_mean=m0, L0_g_prec=P0, L0_g_logconst=c0,
        L1_g_mean=m1, L1_g_prec=P1, L1_g_logconst=c1,
        theta0=np.tile([-8., -7.], (nChains, 1)), z=z, u_c=u_c, u_f=u_f,
        traj=traj, accepted=acc, logpost_L0=lpc, logpost_L1=lpf))


# --------------------------------------------------------------------------
# linear model (C3)
# --------------------------------------------------------------------------

def linear_problem():
    G_f = np.array([[1.4, -0.2], [-0.6, 0.7]])
    b_f = np.zeros(2)
    G_c = G_f + np.array([[-0.6, -0.2], [0.4, 1.1]])
    b_c = np.array([0.5, -0.9])
    truth = np.array([1.5, 0.5])
    rng = Generator(Philox(2222))
    data = np.array([G_f @ truth + b_f + np.sqrt(0.3) * rng.standard_normal(2) for _ in range(5)])
    priorMean = truth + np.array([-0.2, 0.4])
    return dict(G_f=G_f, b_f=b_f, G_c=G_c, b_c=b_c, data=data, noiseVar=np.sqrt(0.3)**2,
                priorMean=priorMean, priorVar=5.0, propVar=0.5)


def linear_models(p):
    data = rh.Data(p['data'])
    noise = rh.CentredGaussianNoise(rh.IIDCovarianceMatrix(2, p['noiseVar']))
    prior = rh.Gaussian(rh.ParameterVector(p['priorMean']), rh.IIDCovarianceMatrix(2, p['priorVar']))
    likC = rh.AdditiveGaussianNoiseLikelihood(data, rh.ForwardModel(rh.LinearSolver(p['G_c'], p['b_c'])), noise)
    likF = rh.AdditiveGaussianNoiseLikelihood(data, rh.ForwardModel(rh.LinearSolver(p['G_f'], p['b_f'])), noise)
    return likC, likF, prior


def linear_level_arrays(p, lvl, prefix):
    return {
        prefix + 'data': p['data'],
        prefix + 'noise_prec': diag_precision(p['noiseVar'], 2),
        prefix + 'prior_mean': p['priorMean'],
        prefix + 'prior_prec': diag_precision(p['priorVar'], 2),
        prefix + 'G': p['G_' + lvl], prefix + 'b': p['b_' + lvl]}


def case_linear(twoLevel):
    from yagremcmc.chain.target import UnnormalisedPosterior
    p = linear_problem()
    nChains, nSteps, J = 3, 300, (5 if twoLevel else 1)
    rng = Generator(Philox(404 + J))
    z, u_c, u_f = make_noise(rng, nChains, nSteps, J, 2,
                             zero_at=[(0, 4, None), (1, 8, 1)] if twoLevel else [(0, 4, 0)])
    traj, acc, lpc, lpf = [], [], [], []
    for c in range(nChains):
        likC, likF, prior = linear_models(p)
        inj = rh.NoiseInjector(z[c], u_c[c], u_f[c])
        if twoLevel:
            b = rh.MLDABuilder()
            b.bayesModel = rh.BayesianRegressionModelHierarchy(
                rh.Hierarchy([likC, likF]), rh.SharedComponent(prior, 2))
            b.baseProposalCovariance = rh.IIDCovarianceMatrix(2, p['propVar'])
            b.subChainLengths = [J]
        else:
            b = rh.MRWBuilder()
            b.bayesModel = rh.BayesianRegressionModel(likF, prior)
            b.proposalCovariance = rh.IIDCovarianceMatrix(2, p['propVar'])
        mcmc = rh.quiet(b.build_method)
        t, a = rh.run_reference_chain(mcmc, rh.ParameterVector(np.zeros(2)), nSteps, inj, twoLevel)
        traj.append(t); acc.append(a)
        lpf.append(logpost_along(UnnormalisedPosterior(likF, prior), rh.ParameterVector, t))
        lpc.append(logpost_along(UnnormalisedPosterior(likC, prior), rh.ParameterVector, t))
    arrays = dict(prop_L=lower_proposal('iid', p['propVar'], 2),
                  theta0=np.zeros((nChains, 2)), z=z, u_c=u_c, u_f=u_f, traj=traj, accepted=acc)
    if twoLevel:
        arrays.update(linear_level_arrays(p, 'c', 'L0_'))
        arrays.update(linear_level_arrays(p, 'f', 'L1_'))
        arrays.update(logpost_L0=lpc, logpost_L1=lpf)
    else:
        arrays.update(linear_level_arrays(p, 'f', 'L0_'))
        arrays.update(logpost_L0=lpf)
    meta = dict(model='linear', dim=2, levels=2 if twoLevel else 1, J=J, eq='exact',
                note='C3 example_inference_linearModel_twoLevel.py:33-74,128-129,173')
    save("mlda_linear" if twoLevel else "mrw_linear", meta, arrays)


def big_linear_problem(d=16, dataDim=24, nData=3, seed=3):
    """GEMM-sized linear model (SURVEY 8d 'GEMM-meaningful variant'): G ~ N(0, 1/d), coarse model = fine + perturbation."""
    rng = Generator(Philox(seed))
    G_f = rng.standard_normal((dataDim, d)) / np.sqrt(d)
    b_f = 0.1 * rng.standard_normal(dataDim)
    G_c = G_f + 0.05 * rng.standard_normal((dataDim, d)) / np.sqrt(d)
    b_c = b_f + 0.02 * rng.standard_normal(dataDim)
    truth = rng.standard_normal(d)
    data = np.array([G_f @ truth + b_f + np.sqrt(0.05) * rng.standard_normal(dataDim) for _ in range(nData)])
    return dict(G_f=G_f, b_f=b_f, G_c=G_c, b_c=b_c, data=data, noiseVar=0.05, priorMean=np.zeros(d) + 0.1,
                priorVar=2.0, propVar=0.004, truth=truth, d=d, dataDim=dataDim)


def case_big_linear(twoLevel, name=None, **pkw):
    from yagremcmc.chain.target import UnnormalisedPosterior
    p = big_linear_problem(**pkw)
    d, dd = p['d'], p['dataDim']
    nChains, nSteps, J = 3, 150, (2 if twoLevel else 1)
    rng = Generator(Philox(1300 + J + 7 * len(pkw)))
    z, u_c, u_f = make_noise(rng, nChains, nSteps, J, d, zero_at=[(0, 4, None)] if twoLevel else [(0, 4, 0)])
    theta0 = p['truth'] + 0.05 * rng.standard_normal((nChains, d))
    traj, acc, lpc, lpf = [], [], [], []
    for c in range(nChains):
        data = rh.Data(p['data'])
        noise = rh.CentredGaussianNoise(rh.IIDCovarianceMatrix(dd, p['noiseVar']))
        prior = rh.Gaussian(rh.ParameterVector(p['priorMean']), rh.IIDCovarianceMatrix(d, p['priorVar']))
        likC = rh.AdditiveGaussianNoiseLikelihood(data, rh.ForwardModel(rh.LinearSolver(p['G_c'], p['b_c'])), noise)
        likF = rh.AdditiveGaussianNoiseLikelihood(data, rh.ForwardModel(rh.LinearSolver(p['G_f'], p['b_f'])), noise)
        inj = rh.NoiseInjector(z[c], u_c[c], u_f[c])
        if twoLevel:
            b = rh.MLDABuilder()
            b.bayesModel = rh.BayesianRegressionModelHierarchy(rh.Hierarchy([likC, likF]), rh.SharedComponent(prior, 2))
            b.baseProposalCovariance = rh.IIDCovarianceMatrix(d, p['propVar'])
            b.subChainLengths = [J]
        else:
            b = rh.MRWBuilder()
            b.bayesModel = rh.BayesianRegressionModel(likF, prior)
            b.proposalCovariance = rh.IIDCovarianceMatrix(d, p['propVar'])
        mcmc = rh.quiet(b.build_method)
        t, a = rh.run_reference_chain(mcmc, rh.ParameterVector(theta0[c].copy()), nSteps, inj, twoLevel)
        traj.append(t); acc.append(a)
        lpf.append(logpost_along(UnnormalisedPosterior(likF, prior), rh.ParameterVector, t))
        lpc.append(logpost_along(UnnormalisedPosterior(likC, prior), rh.ParameterVector, t))
        print(f"    big linear ({'two' if twoLevel else 'single'} level) chain {c}: acceptance {a.mean():.3f}")

    def level(lvl, prefix):
        return {prefix + 'data': p['data'], prefix + 'noise_prec': diag_precision(p['noiseVar'], dd),
                prefix + 'prior_mean': p['priorMean'], prefix + 'prior_prec': diag_precision(p['priorVar'], d),
                prefix + 'G': p['G_' + lvl], prefix + 'b': p['b_' + lvl]}
    arrays = dict(prop_L=lower_proposal('iid', p['propVar'], d), theta0=theta0, z=z, u_c=u_c, u_f=u_f,
                  traj=traj, accepted=acc)
    if twoLevel:
        arrays.update(level('c', 'L0_')); arrays.update(level('f', 'L1_'))
        arrays.update(logpost_L0=lpc, logpost_L1=lpf)
    else:
        arrays.update(level('f', 'L0_'))
        arrays.update(logpost_L0=lpf)
    meta = dict(model='linear', dim=d, levels=2 if twoLevel else 1, J=J, eq='exact',
                note=f'GEMM-sized linear model d={d}, dataDim={dd}, nData={p["data"].shape[0]} (reference chain stack + exampleSetup-style A@theta+b)')
    save(name or ("mlda_linear_big" if twoLevel else "mrw_linear_big"), meta, arrays)


def case_pcn_big_linear():
    """pCN (chain/method/pcn.py) on the GEMM-sized linear model: diagonal centred Gaussian prior, likelihood-only target."""
    p = big_linear_problem(d=12, dataDim=20, nData=2, seed=5)
    d, dd = p['d'], p['dataDim']
    nChains, nSteps, step = 3, 150, 0.002
    rng = Generator(Philox(1390))
    z, u_c, u_f = make_noise(rng, nChains, nSteps, 1, d, zero_at=[(0, 6, 0)])
    priorVar = np.linspace(0.5, 2.0, d)
    theta0 = p['truth'] + 0.05 * rng.standard_normal((nChains, d))
    traj, acc, lp = [], [], []
    for c in range(nChains):
        data = rh.Data(p['data'])
        noise = rh.CentredGaussianNoise(rh.IIDCovarianceMatrix(dd, p['noiseVar']))
        lik = rh.AdditiveGaussianNoiseLikelihood(data, rh.ForwardModel(rh.LinearSolver(p['G_f'], p['b_f'])), noise)
        prior = rh.Gaussian(rh.ParameterVector(np.zeros(d)), rh.DiagonalCovarianceMatrix(priorVar))
        inj = rh.NoiseInjector(z[c], None, u_f[c])
        b = rh.PCNBuilder()
        b.bayesModel = rh.BayesianRegressionModel(lik, prior)
        b.stepSize = step
        mcmc = rh.quiet(b.build_method)
        t, a = rh.run_reference_chain(mcmc, rh.ParameterVector(theta0[c].copy()), nSteps, inj, False)
        traj.append(t); acc.append(a)
        lp.append(logpost_along(lik, rh.ParameterVector, t))
        print(f"    pcn_linear_big chain {c}: acceptance {a.mean():.3f}")
    arrays = dict(prop_L=lower_proposal('diag', priorVar, d), pcn_mean=np.zeros(d), theta0=theta0, z=z, u_c=u_c, u_f=u_f,
                  traj=traj, accepted=acc, logpost_L0=lp,
                  L0_data=p['data'], L0_noise_prec=diag_precision(p['noiseVar'], dd), L0_prior_mean=np.zeros(d),
                  L0_prior_prec=np.zeros((d, d)), L0_G=p['G_f'], L0_b=p['b_f'])
    meta = dict(model='linear', dim=d, levels=1, J=1, eq='exact', proposal='pcn', pcn_step=step,
                note='pCN on a GEMM-sized linear model (d=12, dataDim=20, nData=2), diagonal centred Gaussian prior')
    save("pcn_linear_big", meta, arrays)


def _spd(rng, d, scale):
    A = rng.standard_normal((d, d))
    return scale * (A @ A.T / d + 0.5 * np.eye(d))


def case_big_linear_dense():
    """GEMM-sized linear model with DenseCovarianceMatrix objects (statistics/covariance.py:69-94) as the proposal
    covariance and as the prior covariance: MRW, and pCN (whose proposal factor is the dense prior's)."""
    from yagremcmc.chain.target import UnnormalisedPosterior
    p = big_linear_problem(d=12, dataDim=20, nData=2, seed=6)
    d, dd = p['d'], p['dataDim']
    rng = Generator(Philox(1395))
    propCov, priorCov = _spd(rng, d, 0.004), _spd(rng, d, 1.5)
    nChains, nSteps = 3, 150
    for kind in ('mrw', 'pcn'):
        z, u_c, u_f = make_noise(rng, nChains, nSteps, 1, d, zero_at=[(0, 6, 0)])
        theta0 = p['truth'] + 0.05 * rng.standard_normal((nChains, d))
        traj, acc, lp = [], [], []
        for c in range(nChains):
            data = rh.Data(p['data'])
            noise = rh.CentredGaussianNoise(rh.IIDCovarianceMatrix(dd, p['noiseVar']))
            lik = rh.AdditiveGaussianNoiseLikelihood(data, rh.ForwardModel(rh.LinearSolver(p['G_f'], p['b_f'])), noise)
            priorMean = np.zeros(d) if kind == 'pcn' else p['priorMean']
            prior = rh.Gaussian(rh.ParameterVector(priorMean), rh.DenseCovarianceMatrix(priorCov))
            inj = rh.NoiseInjector(z[c], None, u_f[c])
            if kind == 'mrw':
                b = rh.MRWBuilder()
                b.proposalCovariance = rh.DenseCovarianceMatrix(propCov)
            else:
                b = rh.PCNBuilder()
                b.stepSize = 0.002
            b.bayesModel = rh.BayesianRegressionModel(lik, prior)
            mcmc = rh.quiet(b.build_method)
            t, a = rh.run_reference_chain(mcmc, rh.ParameterVector(theta0[c].copy()), nSteps, inj, False)
            traj.append(t); acc.append(a)
            tgt = lik if kind == 'pcn' else UnnormalisedPosterior(lik, prior)
            lp.append(logpost_along(tgt, rh.ParameterVector, t))
            print(f"    {kind}_linear_big_dense chain {c}: acceptance {a.mean():.3f}")
        prec = np.linalg.inv(priorCov)
        arrays = dict(theta0=theta0, z=z, u_c=u_c, u_f=u_f, traj=traj, accepted=acc, logpost_L0=lp,
                      L0_data=p['data'], L0_noise_prec=diag_precision(p['noiseVar'], dd), L0_G=p['G_f'], L0_b=p['b_f'])
        if kind == 'mrw':
            arrays.update(prop_L=lower_proposal('dense', propCov, d), L0_prior_mean=p['priorMean'],
                          L0_prior_prec=0.5 * (prec + prec.T))
            meta = dict(model='linear', dim=d, levels=1, J=1, eq='exact',
                        note='GEMM-sized linear model (d=12, dataDim=20, nData=2), dense proposal covariance and dense Gaussian prior')
        else:
            arrays.update(prop_L=lower_proposal('dense', priorCov, d), pcn_mean=np.zeros(d), L0_prior_mean=np.zeros(d),
                          L0_prior_prec=np.zeros((d, d)))
            meta = dict(model='linear', dim=d, levels=1, J=1, eq='exact', proposal='pcn', pcn_step=0.002,
                        note='pCN on a GEMM-sized linear model with a dense centred Gaussian prior')
        save(f"{kind}_linear_big_dense", meta, arrays)


# --------------------------------------------------------------------------
# Lotka-Volterra (C4 / C5)
# --------------------------------------------------------------------------

def lv_problem(Nc=64, Nf=512, T=10., nData=10, seed=1112):
    rng = Generator(Philox(seed))
    design = rng.uniform(0.5, 1.5, (nData, 2))
    truth = np.log(np.array([0.4, 0.6]))
    cfg = dict(T=T, alpha=0.8, gamma=0.4, nData=nData, dataDim=2)
    sol = rh.RK4LotkaVolterraSolver(design, dict(cfg, rk4Steps=Nf))
    sol.interpolate(rh.LotkaVolterraParameter.from_coefficient(truth))
    sol.invoke()
    data = sol.evaluation + np.sqrt(0.04) * rng.standard_normal((nData, 2))
    return dict(design=design, truth=truth, cfg=cfg, data=data, Nc=Nc, Nf=Nf,
                noiseVar=0.04, priorVar=1.4, priorMean=np.zeros(2))


def lv_models(p):
    data = rh.Data(p['data'])
    noise = rh.CentredGaussianNoise(rh.IIDCovarianceMatrix(2, p['noiseVar']))
    prior = rh.Gaussian(rh.LotkaVolterraParameter.from_coefficient(p['priorMean']),
                        rh.IIDCovarianceMatrix(2, p['priorVar']))
    solC = rh.RK4LotkaVolterraSolver(p['design'], dict(p['cfg'], rk4Steps=p['Nc']))
    solF = rh.RK4LotkaVolterraSolver(p['design'], dict(p['cfg'], rk4Steps=p['Nf']))
    likC = rh.AdditiveGaussianNoiseLikelihood(data, rh.ForwardModel(solC), noise)
    likF = rh.AdditiveGaussianNoiseLikelihood(data, rh.ForwardModel(solF), noise)
    return likC, likF, prior


def lv_level_arrays(p, N, prefix):
    return {
        prefix + 'data': p['data'],
        prefix + 'noise_prec': diag_precision(p['noiseVar'], 2),
        prefix + 'prior_mean': p['priorMean'],
        prefix + 'prior_prec': diag_precision(p['priorVar'], 2),
        prefix + 'design': p['design'],
        prefix + 'lv': np.array([p['cfg']['alpha'], p['cfg']['gamma'], p['cfg']['T'], float(N)])}


def case_lv(name, twoLevel, p, propVar, nChains, nSteps, J, seed, theta0, zero_at=(), zscale=None, note=''):
    from yagremcmc.chain.target import UnnormalisedPosterior
    rng = Generator(Philox(seed))
    z, u_c, u_f = make_noise(rng, nChains, nSteps, J, 2, zero_at=zero_at)
    if zscale is not None:
        for (c, n, j, s) in zscale:
            z[c, n, j, :] *= s
    traj, acc, lpc, lpf = [], [], [], []
    for c in range(nChains):
        likC, likF, prior = lv_models(p)
        inj = rh.NoiseInjector(z[c], u_c[c], u_f[c])
        if twoLevel:
            b = rh.MLDABuilder()
            b.bayesModel = rh.BayesianRegressionModelHierarchy(
                rh.Hierarchy([likC, likF]), rh.SharedComponent(prior, 2))
            b.baseProposalCovariance = rh.IIDCovarianceMatrix(2, propVar)
            b.subChainLengths = [J]
        else:
            b = rh.MRWBuilder()
            b.bayesModel = rh.BayesianRegressionModel(likF, prior)
            b.proposalCovariance = rh.IIDCovarianceMatrix(2, propVar)
        mcmc = rh.quiet(b.build_method)
        init = rh.LotkaVolterraParameter.from_coefficient(theta0[c].copy())
        t, a = rh.run_reference_chain(mcmc, init, nSteps, inj, twoLevel)
        traj.append(t); acc.append(a)
        lpf.append(logpost_along(UnnormalisedPosterior(likF, prior), rh.LotkaVolterraParameter, t))
        if twoLevel:
            lpc.append(logpost_along(UnnormalisedPosterior(likC, prior), rh.LotkaVolterraParameter, t))
        print(f"    {name} chain {c}: acceptance {a.mean():.3f}")
    arrays = dict(prop_L=lower_proposal('iid', propVar, 2),
                  theta0=theta0, z=z, u_c=u_c, u_f=u_f, traj=traj, accepted=acc)
    if twoLevel:
        arrays.update(lv_level_arrays(p, p['Nc'], 'L0_'))
        arrays.update(lv_level_arrays(p, p['Nf'], 'L1_'))
        arrays.update(logpost_L0=lpc, logpost_L1=lpf)
    else:
        arrays.update(lv_level_arrays(p, p['Nf'], 'L0_'))
        arrays.update(logpost_L0=lpf)
    meta = dict(model='lv', dim=2, levels=2 if twoLevel else 1, J=J, eq='exact', note=note)
    save(name, meta, arrays)


def cases_lv():
    p = lv_problem()
    rng = Generator(Philox(7))
    theta0 = p['truth'] + 0.05 * rng.standard_normal((3, 2))
    case_lv("mrw_lv", False, p, 0.15, 3, 120, 1, 505, theta0, zero_at=[(0, 6, 0)],
            note='C4 example_inference_lotkaVolterra_singleLevel.py:29-59,82-83 with RK4 N=512')
    case_lv("mlda_lv", True, p, 0.1, 3, 120, 3, 606, theta0,
            zero_at=[(0, 4, None), (1, 9, 1)],
            note='C5 example_inference_lotkaVolterra_twoLevel.py:29-44,95-106 with RK4 Nc=64 Nf=512, J=3')
    # non-finite forward outputs: tiny RK4 step count + huge proposals, and a start in the NaN region
    pe = lv_problem(Nc=8, Nf=16)
    theta0e = np.array([[-7., 2.8], p['truth'], [3.0, 4.0]])
    case_lv("mlda_lv_nonfinite", True, pe, 0.1, 3, 60, 2, 707, theta0e,
            zscale=[(1, 3, 0, 40.0), (1, 10, 1, 60.0), (0, 2, 0, 25.0), (2, 5, 0, 30.0)],
            note='edge: RK4 overflow -> +inf forward output -> logL=-inf; -inf/-inf -> NaN ratio accepts '
                 '(mrw.py:54-57); SURVEY section 7')


# --------------------------------------------------------------------------
# preconditioned Crank-Nicolson (chain/method/pcn.py; used by
# test/test_inference_mcmc_singleLevel.py:121-148 and example_inference_lotkaVolterra_singleLevel.py:73-76)
# --------------------------------------------------------------------------

def case_pcn(name, model, stepSize, nChains, nSteps, seed, theta0, priorKind, priorValue, note):
    rng = Generator(Philox(seed))
    z, u_c, u_f = make_noise(rng, nChains, nSteps, 1, 2, zero_at=[(0, 6, 0)])
    if model == 'lv':
        p = lv_problem()
        stateType = rh.LotkaVolterraParameter
    else:
        p = linear_problem()
        stateType = rh.ParameterVector
    traj, acc, lp = [], [], []
    for c in range(nChains):
        if model == 'lv':
            _, lik, _ = lv_models(p)
        else:
            _, lik, _ = linear_models(p)
        prior = rh.Gaussian(stateType(np.zeros(2)), rh.covariance_from_spec(priorKind, priorValue, 2))
        inj = rh.NoiseInjector(z[c], None, u_f[c])
        b = rh.PCNBuilder()
        b.bayesModel = rh.BayesianRegressionModel(lik, prior)
        b.stepSize = stepSize
        mcmc = rh.quiet(b.build_method)
        t, a = rh.run_reference_chain(mcmc, stateType(theta0[c].copy()), nSteps, inj, False)
        traj.append(t); acc.append(a)
        lp.append(logpost_along(lik, stateType, t))          # the pCN target is the likelihood alone (pcn.py:52-57)
        print(f"    {name} chain {c}: acceptance {a.mean():.3f}")
    arrays = dict(prop_L=lower_proposal(priorKind, priorValue, 2), pcn_mean=np.zeros(2),
                  theta0=theta0, z=z, u_c=u_c, u_f=u_f, traj=traj, accepted=acc, logpost_L0=lp)
    lvl = lv_level_arrays(p, p['Nf'], 'L0_') if model == 'lv' else linear_level_arrays(p, 'f', 'L0_')
    lvl['L0_prior_mean'] = np.zeros(2)
    lvl['L0_prior_prec'] = np.zeros((2, 2))                  # likelihood-only target
    arrays.update(lvl)
    meta = dict(model=model, dim=2, levels=1, J=1, eq='exact', proposal='pcn', pcn_step=stepSize, note=note)
    save(name, meta, arrays)


def cases_pcn():
    p = lv_problem()
    rng = Generator(Philox(9))
    theta0 = p['truth'] + 0.05 * rng.standard_normal((3, 2))
    case_pcn("pcn_lv", 'lv', 0.004, 3, 120, 909, theta0, 'iid', 1.4,
             'pCN on the C4 likelihood (RK4 N=512), prior N(0, 1.4 I), step 0.004')
    case_pcn("pcn_linear_dense", 'linear', 0.05, 3, 300, 910, np.tile([1.0, 0.3], (3, 1)), 'dense',
             [[2.0, 0.6], [0.6, 1.0]], 'pCN on the C3 fine likelihood, dense centred Gaussian prior, step 0.05')


# --------------------------------------------------------------------------
# adaptive error model (chain/method/aem.py, statistics/likelihood.py:90-155, statistics/noise.py:25-61;
# example_inference_linearModel_twoLevel.py:95-102,183-191)
# --------------------------------------------------------------------------

def case_aem(name, minDataSize, useHeuristic, nSteps, seed):
    p = linear_problem()
    nChains, J = 3, 5
    rng = Generator(Philox(seed))
    z, u_c, u_f = make_noise(rng, nChains, nSteps, J, 2, zero_at=[(0, 4, None), (1, 8, 1)])
    traj, acc, en, em, ev, nev = [], [], [], [], [], []
    for c in range(nChains):
        data = rh.Data(p['data'])
        noise = rh.CentredGaussianNoise(rh.IIDCovarianceMatrix(2, p['noiseVar']))
        prior = rh.Gaussian(rh.ParameterVector(p['priorMean']), rh.IIDCovarianceMatrix(2, p['priorVar']))
        likC = rh.AEMLikelihood(data, rh.ForwardModel(rh.LinearSolver(p['G_c'], p['b_c'])), noise, minDataSize, useHeuristic)
        likF = rh.AEMLikelihood(data, rh.ForwardModel(rh.LinearSolver(p['G_f'], p['b_f'])), noise, minDataSize, useHeuristic)
        inj = rh.NoiseInjector(z[c], u_c[c], u_f[c])
        b = rh.AEMBuilder()
        b.bayesModel = rh.BayesianRegressionModelHierarchy(rh.Hierarchy([likC, likF]), rh.SharedComponent(prior, 2))
        b.baseProposalCovariance = rh.IIDCovarianceMatrix(2, p['propVar'])
        b.subChainLengths = [J]
        mcmc = rh.quiet(b.build_method)
        t, a = rh.run_reference_chain(mcmc, rh.ParameterVector(np.zeros(2)), nSteps, inj, True)
        traj.append(t); acc.append(a)
        accum = likC.accumulator
        en.append(accum.nData)
        em.append(accum.mean() if accum.nData > 0 else np.zeros(2))
        ev.append(accum.marginal_variance() if accum.nData > 1 else np.zeros(2))
        nev.append([likC.number_of_model_evaluations(), likF.number_of_model_evaluations()])
        print(f"    {name} chain {c}: acceptance {a.mean():.3f}, error samples {accum.nData}, "
              f"model evaluations {nev[-1]}")
    arrays = dict(prop_L=lower_proposal('iid', p['propVar'], 2), theta0=np.zeros((nChains, 2)),
                  z=z, u_c=u_c, u_f=u_f, traj=traj, accepted=acc,
                  aem_n=np.array(en), aem_mean=np.array(em), aem_var=np.array(ev), n_model_evals=np.array(nev))
    arrays.update(linear_level_arrays(p, 'c', 'L0_'))
    arrays.update(linear_level_arrays(p, 'f', 'L1_'))
    meta = dict(model='linear', dim=2, levels=2, J=J, eq='exact', aem=dict(min_data=minDataSize, heuristic=bool(useHeuristic)),
                note='AEM on C3: example_inference_linearModel_twoLevel.py:95-102,183-191')
    save(name, meta, arrays)


def cases_aem():
    case_aem("aem_linear", 5, True, 400, 1500)
    case_aem("aem_linear_noheuristic", 12, False, 400, 1501)



# --------------------------------------------------------------------------
# adaptive Metropolis through the reference's live interface (chain/adaptive.py:8-64):
# the unmodified AdaptiveMRWProposal + MetropolisHastings.run drive OUR concrete
# AdaptiveCovarianceMatrix (ref_harness.HaarioAdaptiveCovariance)
# --------------------------------------------------------------------------

def _am_finish(name, meta, arrays, covs, am):
    arrays.update(am_mean=np.array([c.mean for c in covs]), am_m2=np.array([c.M2 for c in covs]),
                  am_L=np.array([c.chol_factor() for c in covs]), am_t=np.array([c.t for c in covs]),
                  am_refreshes=np.array([c.nRefresh for c in covs]))
    meta['am'] = am
    for c, cv in enumerate(covs):
        print(f"    {name} chain {c}: {cv.t} updates, {cv.nRefresh} Cholesky refreshes")
    save(name, meta, arrays)


def case_am_gauss2d(name, am, seed, nSteps=300):
    nChains = 3
    rng = Generator(Philox(seed))
    z, u_c, u_f = make_noise(rng, nChains, nSteps, 1, 2, zero_at=[(2, 11, 0)])
    mean, cov = np.array([1., 1.5]), np.array([[3.2, -0.4], [-0.4, 0.2]])       # test/test_adaptive.py:16-19
    tgt = rh.GaussianTargetDensity2d(rh.ParameterVector(mean), cov)
    traj, acc, lp, covs = [], [], [], []
    for c in range(nChains):
        inj = rh.NoiseInjector(z[c], None, u_f[c])
        ac = rh.HaarioAdaptiveCovariance(rh.IIDCovarianceMatrix(2, 0.25), am['idle'], am['collection'], am['eps'],
                                         am.get('scale'), am.get('refresh', 1))
        mcmc = rh.AdaptiveMetropolis(tgt, ac, rh.FullDiagnostics())
        t, a = rh.run_reference_chain(mcmc, rh.ParameterVector(np.array([0., -1.])), nSteps, inj, False)
        traj.append(t); acc.append(a); covs.append(ac)
        lp.append(logpost_along(tgt, rh.ParameterVector, t))
    m, P, lc = gauss2d_level_arrays(mean, cov)
    meta = dict(model='gauss', dim=2, levels=1, J=1, eq='exact',
                note='adaptive Metropolis: reference AdaptiveMRWProposal (chain/adaptive.py:37-64) + '
                     'MetropolisHastings.run + our HaarioAdaptiveCovariance; target of test/test_adaptive.py:16-21')
    _am_finish(name, meta, dict(prop_L=lower_proposal('iid', 0.25, 2), L0_g_mean=m, L0_g_prec=P, L0_g_logconst=lc,
                                theta0=np.tile([0., -1.], (nChains, 1)), z=z, u_c=u_c, u_f=u_f, traj=traj,
                                accepted=acc, logpost_L0=lp), covs, am)


def case_am_lv(name, twoLevel, am, seed, nSteps=120, J=3):
    from yagremcmc.chain.target import UnnormalisedPosterior
    p = lv_problem()
    nChains = 3
    J = J if twoLevel else 1
    rng = Generator(Philox(seed))
    z, u_c, u_f = make_noise(rng, nChains, nSteps, J, 2, zero_at=[(0, 4, None), (1, 9, 1)] if twoLevel else [(0, 6, 0)])
    theta0 = p['truth'] + 0.05 * Generator(Philox(7)).standard_normal((3, 2))
    propVar = 0.1 if twoLevel else 0.15
    traj, acc, lpc, lpf, covs = [], [], [], [], []
    for c in range(nChains):
        likC, likF, prior = lv_models(p)
        inj = rh.NoiseInjector(z[c], u_c[c], u_f[c])
        ac = rh.HaarioAdaptiveCovariance(rh.IIDCovarianceMatrix(2, propVar), am['idle'], am['collection'], am['eps'],
                                         am.get('scale'), am.get('refresh', 1))
        if twoLevel:
            b = rh.MLDABuilder()
            b.bayesModel = rh.BayesianRegressionModelHierarchy(
                rh.Hierarchy([likC, likF]), rh.SharedComponent(prior, 2))
            b.baseProposalCovariance = rh.IIDCovarianceMatrix(2, propVar)
            b.subChainLengths = [J]
            mcmc = rh.make_surrogate_adaptive(rh.quiet(b.build_method), ac)
        else:
            mcmc = rh.AdaptiveMetropolis(UnnormalisedPosterior(likF, prior), ac, rh.FullDiagnostics())
        init = rh.LotkaVolterraParameter.from_coefficient(theta0[c].copy())
        t, a = rh.run_reference_chain(mcmc, init, nSteps, inj, twoLevel)
        traj.append(t); acc.append(a); covs.append(ac)
        lpf.append(logpost_along(UnnormalisedPosterior(likF, prior), rh.LotkaVolterraParameter, t))
        if twoLevel:
            lpc.append(logpost_along(UnnormalisedPosterior(likC, prior), rh.LotkaVolterraParameter, t))
        print(f"    {name} chain {c}: acceptance {a.mean():.3f}")
    arrays = dict(prop_L=lower_proposal('iid', propVar, 2), theta0=theta0, z=z, u_c=u_c, u_f=u_f, traj=traj, accepted=acc)
    if twoLevel:
        arrays.update(lv_level_arrays(p, p['Nc'], 'L0_')); arrays.update(lv_level_arrays(p, p['Nf'], 'L1_'))
        arrays.update(logpost_L0=lpc, logpost_L1=lpf)
    else:
        arrays.update(lv_level_arrays(p, p['Nf'], 'L0_'))
        arrays.update(logpost_L0=lpf)
    meta = dict(model='lv', dim=2, levels=2 if twoLevel else 1, J=J, eq='exact',
                note=('C5' if twoLevel else 'C4') + ' with an adaptive (coarse) proposal: reference AdaptiveMRWProposal '
                     '(chain/adaptive.py:37-64) ' + ('as the proposal method of the MLDA surrogate MRW, update() before '
                     'every coarse proposal' if twoLevel else '+ MetropolisHastings.run'))
    _am_finish(name, meta, arrays, covs, am)


def case_am_mlda_linear(name, am, seed, nSteps=300, J=5):
    from yagremcmc.chain.target import UnnormalisedPosterior
    p = linear_problem()
    nChains = 3
    rng = Generator(Philox(seed))
    z, u_c, u_f = make_noise(rng, nChains, nSteps, J, 2, zero_at=[(0, 4, None), (1, 8, 1)])
    traj, acc, lpc, lpf, covs = [], [], [], [], []
    for c in range(nChains):
        likC, likF, prior = linear_models(p)
        inj = rh.NoiseInjector(z[c], u_c[c], u_f[c])
        ac = rh.HaarioAdaptiveCovariance(rh.IIDCovarianceMatrix(2, p['propVar']), am['idle'], am['collection'],
                                         am['eps'], am.get('scale'), am.get('refresh', 1))
        b = rh.MLDABuilder()
        b.bayesModel = rh.BayesianRegressionModelHierarchy(rh.Hierarchy([likC, likF]), rh.SharedComponent(prior, 2))
        b.baseProposalCovariance = rh.IIDCovarianceMatrix(2, p['propVar'])
        b.subChainLengths = [J]
        mcmc = rh.make_surrogate_adaptive(rh.quiet(b.build_method), ac)
        t, a = rh.run_reference_chain(mcmc, rh.ParameterVector(np.zeros(2)), nSteps, inj, True)
        traj.append(t); acc.append(a); covs.append(ac)
        lpf.append(logpost_along(UnnormalisedPosterior(likF, prior), rh.ParameterVector, t))
        lpc.append(logpost_along(UnnormalisedPosterior(likC, prior), rh.ParameterVector, t))
        print(f"    {name} chain {c}: acceptance {a.mean():.3f}")
    arrays = dict(prop_L=lower_proposal('iid', p['propVar'], 2), theta0=np.zeros((nChains, 2)), z=z, u_c=u_c, u_f=u_f,
                  traj=traj, accepted=acc, logpost_L0=lpc, logpost_L1=lpf)
    arrays.update(linear_level_arrays(p, 'c', 'L0_')); arrays.update(linear_level_arrays(p, 'f', 'L1_'))
    meta = dict(model='linear', dim=2, levels=2, J=J, eq='exact',
                note='C3 with an adaptive coarse proposal (AdaptiveMRWProposal on the MLDA surrogate MRW)')
    _am_finish(name, meta, arrays, covs, am)


def cases_am():
    case_am_gauss2d("am_gauss2d", dict(idle=20, collection=40, eps=1e-4, refresh=1), 1601)
    case_am_gauss2d("am_gauss2d_refresh7", dict(idle=0, collection=25, eps=1e-3, refresh=7, scale=1.7), 1602)
    case_am_lv("am_lv", False, dict(idle=10, collection=30, eps=1e-6, refresh=1), 1603)
    case_am_lv("am_mlda_lv", True, dict(idle=15, collection=60, eps=1e-6, refresh=1), 1604)
    case_am_mlda_linear("am_mlda_linear", dict(idle=10, collection=100, eps=1e-4, refresh=3), 1605)


# --------------------------------------------------------------------------
# MLDA with TWO surrogates, as the reference actually runs it (chain/method/mlda.py:12-43,60-71,112-117):
# the proposal is the end point of subChainLengths[1] MRW steps on the BASE surrogate (subChainLengths[0]
# is ignored) and the screen uses the FINEST surrogate.  Constants: test/test_mlda.py:14-60 and
# example_mcmc_2d_hierarchical.py:10-44.
# --------------------------------------------------------------------------

def case_mlda3_gauss2d(name, baseCovScale, fineCovScale, nStepsList, seed, nSteps=250, distinctLengths=False):
    nChains = 3
    J = nStepsList[1]
    rng = Generator(Philox(seed))
    z, u_c, u_f = make_noise(rng, nChains, nSteps, J, 2, zero_at=[(0, 3, None), (1, 14, 2), (2, 30, None)])
    baseMean, fineMean = TGT_MEAN + np.array([-0.05, 0.01]), TGT_MEAN + np.array([0.0, -0.01])
    baseCov = baseCovScale * np.array([[2.8, -0.1], [-0.1, 1.7]])
    fineCov = fineCovScale * np.array([[2.4, -0.3], [-0.3, 1.1]])
    tgt = rh.GaussianTargetDensity2d(rh.ParameterVector(TGT_MEAN), TGT_COV)
    base = rh.GaussianTargetDensity2d(rh.ParameterVector(baseMean), baseCov)
    fine = rh.GaussianTargetDensity2d(rh.ParameterVector(fineMean), fineCov)
    traj, acc, lp0, lp1, lp2, order = [], [], [], [], [], []
    for c in range(nChains):
        inj = rh.NoiseInjector(z[c], u_c[c], u_f[c])
        b = rh.MLDABuilder()
        b.explicitTarget = tgt
        b.surrogateTargets = [base, fine]
        b.baseProposalCovariance = rh.IIDCovarianceMatrix(2, 1.)
        b.subChainLengths = list(nStepsList)
        mcmc = rh.quiet(b.build_method)
        assert mcmc.nSurrogates == 2
        t, a = rh.run_reference_chain(mcmc, rh.ParameterVector(np.array([-8., -7.])), nSteps, inj, True)
        traj.append(t); acc.append(a)
        lp0.append(logpost_along(base, rh.ParameterVector, t))
        lp1.append(logpost_along(fine, rh.ParameterVector, t))
        lp2.append(logpost_along(tgt, rh.ParameterVector, t))
        order.append(''.join(inj.log))
        print(f"    {name} chain {c}: acceptance {a.mean():.3f}")
    m0, P0, c0 = gauss2d_level_arrays(baseMean, baseCov)
    m1, P1, c1 = gauss2d_level_arrays(fineMean, fineCov)
    m2, P2, c2 = gauss2d_level_arrays(TGT_MEAN, TGT_COV)
    meta = dict(model='gauss', dim=2, levels=3, J=J, eq='exact', rng_order=order, subChainLengths=list(nStepsList),
                note='MLDA with two surrogates (mlda.py:12-43,60-71,112-117): test/test_mlda.py:14-60 / '
                     'example_mcmc_2d_hierarchical.py:10-44')
    save(name, meta, dict(prop_L=lower_proposal('iid', 1., 2),
                          L0_g_mean=m0, L0_g_prec=P0, L0_g_logconst=c0, L1_g_mean=m1, L1_g_prec=P1, L1_g_logconst=c1,
                          L2_g_mean=m2, L2_g_prec=P2, L2_g_logconst=c2,
                          theta0=np.tile([-8., -7.], (nChains, 1)), z=z, u_c=u_c, u_f=u_f,
                          traj=traj, accepted=acc, logpost_L0=lp0, logpost_L1=lp1, logpost_L2=lp2))


def case_mlda3_linear():
    """Three linear models through the Bayesian-model hierarchy (BayesianRegressionModelHierarchy of size 3)."""
    from yagremcmc.chain.target import UnnormalisedPosterior
    p = linear_problem()
    G_m = p['G_f'] + 0.3 * np.array([[-0.6, -0.2], [0.4, 1.1]])
    b_m = 0.3 * p['b_c']
    nChains, nSteps, J = 3, 250, 4
    rng = Generator(Philox(1707))
    z, u_c, u_f = make_noise(rng, nChains, nSteps, J, 2, zero_at=[(0, 4, None), (1, 8, 1)])
    traj, acc, lps = [], [], [[], [], []]
    for c in range(nChains):
        data = rh.Data(p['data'])
        noise = rh.CentredGaussianNoise(rh.IIDCovarianceMatrix(2, p['noiseVar']))
        prior = rh.Gaussian(rh.ParameterVector(p['priorMean']), rh.IIDCovarianceMatrix(2, p['priorVar']))
        liks = [rh.AdditiveGaussianNoiseLikelihood(data, rh.ForwardModel(rh.LinearSolver(G, b)), noise)
                for (G, b) in ((p['G_c'], p['b_c']), (G_m, b_m), (p['G_f'], p['b_f']))]
        inj = rh.NoiseInjector(z[c], u_c[c], u_f[c])
        b = rh.MLDABuilder()
        b.bayesModel = rh.BayesianRegressionModelHierarchy(rh.Hierarchy(liks), rh.SharedComponent(prior, 3))
        b.baseProposalCovariance = rh.IIDCovarianceMatrix(2, p['propVar'])
        b.subChainLengths = [7, J]
        mcmc = rh.quiet(b.build_method)
        t, a = rh.run_reference_chain(mcmc, rh.ParameterVector(np.zeros(2)), nSteps, inj, True)
        traj.append(t); acc.append(a)
        for l in range(3):
            lps[l].append(logpost_along(UnnormalisedPosterior(liks[l], prior), rh.ParameterVector, t))
        print(f"    mlda3_linear chain {c}: acceptance {a.mean():.3f}")
    arrays = dict(prop_L=lower_proposal('iid', p['propVar'], 2), theta0=np.zeros((nChains, 2)), z=z, u_c=u_c, u_f=u_f,
                  traj=traj, accepted=acc, logpost_L0=lps[0], logpost_L1=lps[1], logpost_L2=lps[2])
    for l, (G, bb) in enumerate(((p['G_c'], p['b_c']), (G_m, b_m), (p['G_f'], p['b_f']))):
        pre = f'L{l}_'
        arrays.update({pre + 'data': p['data'], pre + 'noise_prec': diag_precision(p['noiseVar'], 2),
                       pre + 'prior_mean': p['priorMean'], pre + 'prior_prec': diag_precision(p['priorVar'], 2),
                       pre + 'G': G, pre + 'b': bb})
    meta = dict(model='linear', dim=2, levels=3, J=J, eq='exact', subChainLengths=[7, J],
                note='MLDA over a hierarchy of three linear Bayesian models (C3 constants + an intermediate model)')
    save("mlda3_linear", meta, arrays)


def case_mlda_linear_tempered():
    """Two-level delayed acceptance whose surrogate is a TemperedUnnormalisedPosterior (chain/target.py:25-43:
    tempering * logL + logprior) -- the working part of the reference's tempering machinery (tmlda.py itself
    cannot be constructed: TemperedMLDA.__init__ calls MLDA.__init__ with two arguments missing)."""
    from yagremcmc.chain.target import UnnormalisedPosterior, TemperedUnnormalisedPosterior
    p = linear_problem()
    nChains, nSteps, J, tempering = 3, 250, 5, 0.35
    rng = Generator(Philox(1801))
    z, u_c, u_f = make_noise(rng, nChains, nSteps, J, 2, zero_at=[(0, 4, None), (1, 8, 1)])
    traj, acc, lpc, lpf = [], [], [], []
    for c in range(nChains):
        likC, likF, prior = linear_models(p)
        sur, tgt = TemperedUnnormalisedPosterior(likC, prior, tempering), UnnormalisedPosterior(likF, prior)
        inj = rh.NoiseInjector(z[c], u_c[c], u_f[c])
        b = rh.MLDABuilder()
        b.explicitTarget = tgt
        b.surrogateTargets = [sur]
        b.baseProposalCovariance = rh.IIDCovarianceMatrix(2, p['propVar'])
        b.subChainLengths = [J]
        mcmc = rh.quiet(b.build_method)
        t, a = rh.run_reference_chain(mcmc, rh.ParameterVector(np.zeros(2)), nSteps, inj, True)
        traj.append(t); acc.append(a)
        lpf.append(logpost_along(tgt, rh.ParameterVector, t))
        lpc.append(logpost_along(sur, rh.ParameterVector, t))
        print(f"    mlda_linear_tempered chain {c}: acceptance {a.mean():.3f}")
    arrays = dict(prop_L=lower_proposal('iid', p['propVar'], 2), theta0=np.zeros((nChains, 2)), z=z, u_c=u_c, u_f=u_f,
                  traj=traj, accepted=acc, logpost_L0=lpc, logpost_L1=lpf, L0_tempering=tempering)
    arrays.update(linear_level_arrays(p, 'c', 'L0_')); arrays.update(linear_level_arrays(p, 'f', 'L1_'))
    meta = dict(model='linear', dim=2, levels=2, J=J, eq='exact',
                note='C3 with a tempered surrogate: TemperedUnnormalisedPosterior(likC, prior, 0.35), chain/target.py:25-43')
    save("mlda_linear_tempered", meta, arrays)


def cases_mlda3():
    case_mlda3_gauss2d("mlda3_gauss2d", 3.0, 1.5, [6, 6], 1701)                 # test/test_mlda.py:14-60
    case_mlda3_gauss2d("mlda3_gauss2d_hier", 4.0, 2.0, [9, 4], 1702)            # example_mcmc_2d_hierarchical.py ([4,4]);
    case_mlda3_linear()                                                        # [9,4]: subChainLengths[0] is ignored
    case_mlda_linear_tempered()

# --------------------------------------------------------------------------
# post-processing pins: IAT, Welford, dense covariance
# --------------------------------------------------------------------------

def case_postprocessing():
    from yagremcmc.postprocessing.autocorrelation import (
        integrated_autocorrelation, estimate_autocorrelation_function_1d)
    from yagremcmc.statistics.estimation import WelfordAccumulator
    rng = Generator(Philox(808))
    seqs, iat_max, iat_mean, acfs = [], [], [], []
    for rho, n in [(0.0, 500), (0.5, 1000), (0.9, 2000), (0.97, 3000), (0.8, 777)]:
        e = rng.standard_normal((n, 2))
        x = np.zeros((n, 2))
        for i in range(1, n):
            x[i] = np.array([rho, 0.5 * rho]) * x[i - 1] + e[i]
        seqs.append(x)
        iat_max.append(integrated_autocorrelation(x, 'max'))
        iat_mean.append(integrated_autocorrelation(x, 'mean'))
        acfs.append(estimate_autocorrelation_function_1d(x[:, 0])[:64])
    arrays = {f'seq{i}': s for i, s in enumerate(seqs)}
    arrays.update({f'acf{i}': a for i, a in enumerate(acfs)})
    arrays['iat_max'] = np.array(iat_max)
    arrays['iat_mean'] = np.array(iat_mean)
    # Welford (estimation.py:36-53)
    w = WelfordAccumulator()
    xs = rng.standard_normal((1000, 3)) * np.array([1., 3., 0.1]) + np.array([5., -2., 0.])
    for x in xs:
        w.update(x)
    arrays['welford_x'] = xs
    arrays['welford_mean'] = w.mean()
    arrays['welford_var'] = w.marginal_variance()
    arrays['welford_cond'] = w.condition_number()
    # dense covariance operator (covariance.py:69-94)
    C = np.array([[2.0, 0.6, -0.3], [0.6, 1.0, 0.2], [-0.3, 0.2, 0.5]])
    dc = rh.DenseCovarianceMatrix(C)
    v = rng.standard_normal((16, 3))
    arrays['dense_C'] = C
    arrays['dense_v'] = v
    arrays['dense_chol_apply'] = np.array([dc.apply_chol_factor(x) for x in v])
    arrays['dense_inv_apply'] = np.array([dc.apply_inverse(x) for x in v])
    arrays['dense_norm2'] = np.array([dc.induced_norm_squared(x) for x in v])
    save("postprocessing", dict(note='autocorrelation.py:5-140, estimation.py:36-53, covariance.py:69-94'), arrays)

# --------------------------------------------------------------------------
# long seeded runs of the UNMODIFIED reference (its own numpy RNG, no injection): posterior
# moments / acceptance / IAT of C5 and C4, the Monte-Carlo pin north_star asks for; and the
# forward map of the reference's shipped solver (test/testSetup.py:101-141, solve_ivp) at tight
# tolerance, the accuracy pin of the RK4 replacement (style of test/test_solver_invoke.py:64-94)
# --------------------------------------------------------------------------

def _lv_long_worker(args):
    twoLevel, seed, nSteps, burn = args
    import numpy.random as npr
    from yagremcmc.postprocessing.autocorrelation import integrated_autocorrelation
    npr.seed(seed)                      # the reference draws from numpy's legacy global RNG
    p = lv_problem()
    likC, likF, prior = lv_models(p)
    if twoLevel:
        b = rh.MLDABuilder()
        b.bayesModel = rh.BayesianRegressionModelHierarchy(
            rh.Hierarchy([likC, likF]), rh.SharedComponent(prior, 2))
        b.baseProposalCovariance = rh.IIDCovarianceMatrix(2, 0.1)
        b.subChainLengths = [3]
    else:
        b = rh.MRWBuilder()
        b.bayesModel = rh.BayesianRegressionModel(likF, prior)
        b.proposalCovariance = rh.IIDCovarianceMatrix(2, 0.15)
    mcmc = rh.quiet(b.build_method)
    rng = Generator(Philox(7000 + seed))
    init = rh.LotkaVolterraParameter.from_coefficient(p['truth'] + 0.05 * rng.standard_normal(2))
    rh.quiet(mcmc.run, nSteps, init, verbose=False)
    traj = np.array([np.asarray(s, dtype=np.float64).reshape(-1) for s in mcmc.chain.trajectory])
    iat = integrated_autocorrelation(traj[burn:], 'max')
    return traj[burn:], mcmc.diagnostics.global_acceptance_rate(), iat


def case_lv_long(twoLevel, nProc=8, nSteps=9000, burn=200):
    import multiprocessing as mp
    name = "lv_long_twoLevel" if twoLevel else "lv_long_singleLevel"
    with mp.get_context("fork").Pool(nProc) as pool:
        res = pool.map(_lv_long_worker, [(twoLevel, 100 + c, nSteps, burn) for c in range(nProc)])
    chains = np.stack([r[0] for r in res])                       # [nProc, n, 2]
    acc = np.array([r[1] for r in res]); iat = np.array([r[2] for r in res], dtype=np.float64)
    flat = chains.reshape(-1, 2)
    ess = float(sum(chains.shape[1] // max(int(t), 1) for t in iat))
    arrays = dict(chain_means=chains.mean(axis=1), mean=flat.mean(axis=0), cov=np.cov(flat.T),
                  acceptance=acc, iat=iat, ess=np.array(ess), n_kept=np.array(chains.shape[1]))
    meta = dict(model='lv', levels=2 if twoLevel else 1, nChains=nProc, nSteps=nSteps, burnIn=burn,
                note='unmodified reference chain stack + RK4 plugin, numpy legacy RNG seeded per chain; '
                     'problem = lv_problem() (C5 / C4 constants)')
    print(f"    {name}: mean {arrays['mean']}, acceptance {acc.mean():.3f}, IAT {iat}, ESS {ess:.0f}")
    save(name, meta, arrays)


def case_lv_forward():
    """Forward map of the reference's own LotkaVolterraSolver (solve_ivp DOP853 at rtol 1e-12; the class
    does not expose atol, so scipy's default 1e-6 bounds the accuracy) next to the RK4 plugin's, at the
    truth and at a few posterior-scale points."""
    from yagremcmc.test.testSetup import LotkaVolterraSolver
    p = lv_problem()
    cfg = dict(p['cfg'], T=p['cfg']['T'], solver='DOP853', rtol=1e-12)
    thetas = np.stack([p['truth'], p['truth'] + [0.1, -0.08], p['truth'] + [-0.15, 0.12], [-0.5, -0.9]])
    out_ref, out_rk4 = [], {64: [], 512: [], 1024: []}
    for th in thetas:
        prm = rh.LotkaVolterraParameter.from_coefficient(th)
        sol = LotkaVolterraSolver(p['design'], cfg)
        sol.interpolate(prm); sol.invoke()
        out_ref.append(np.array(sol.evaluation))
        for N in out_rk4:
            s2 = rh.RK4LotkaVolterraSolver(p['design'], dict(p['cfg'], rk4Steps=N))
            s2.interpolate(prm); s2.invoke()
            out_rk4[N].append(s2.evaluation)
    arrays = dict(thetas=thetas, design=p['design'], lv=np.array([0.8, 0.4, p['cfg']['T']]),
                  ref_forward=np.stack(out_ref),
                  **{f"rk4_{N}": np.stack(v) for N, v in out_rk4.items()})
    for N in out_rk4:
        e = np.abs(arrays[f"rk4_{N}"] - arrays['ref_forward']) / np.abs(arrays['ref_forward'])
        print(f"    lv_forward: RK4 N={N} vs reference DOP853: max rel err {e.max():.3e}")
    save("lv_forward", dict(model='lv', note='reference LotkaVolterraSolver (testSetup.py:101-141), '
                            'DOP853 rtol 1e-12, vs the RK4 plugin'), arrays)


if __name__ == "__main__":
    only = set(sys.argv[1:])

    def want(k):
        return not only or k in only
    if want('gauss'):
        case_gauss1d()
        case_gauss2d("mrw_gauss2d_iid", 'iid', 1.0, 201)
        case_gauss2d("mrw_gauss2d_diag", 'diag', [1.5, 0.4], 202)
        case_gauss2d("mrw_gauss2d_dense", 'dense', [[1.2, -0.3], [-0.3, 0.5]], 203)
        case_mlda_gauss2d()
    if want('linear'):
        case_linear(False)
        case_linear(True)
    if want('biglinear'):
        case_big_linear(False)
        case_big_linear(True)
    if want('biglinear2'):
        case_big_linear(False, name="mrw_linear_big_rows12", d=10, dataDim=33, nData=12, seed=4)     # more than 8 data rows
        case_pcn_big_linear()
        case_big_linear_dense()
    if want('post'):
        case_postprocessing()
    if want('lv'):
        cases_lv()
    if want('pcn'):
        cases_pcn()
    if want('aem'):
        cases_aem()
    if want('am'):
        cases_am()
    if want('mlda3'):
        cases_mlda3()
    if want('lvforward'):
        case_lv_forward()
    if want('lvlong'):
        case_lv_long(True)
        case_lv_long(False)
