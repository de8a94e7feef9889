"""
TEST / MEASUREMENT INFRASTRUCTURE, CONTAINER-ONLY: times the UNMODIFIED pure-Python reference
(/root/reference) on the C5 problem, the two CPU baselines SURVEY 8d / BASELINE.md section 3 name:

  (i)  like for like: the reference chain stack (MLDABuilder -> MetropolisHastings.run) with the fixed-step
       RK4 SolverInterface plugin of oracle/ref_harness.py (Nc = 64, Nf = 512), and
  (ii) as shipped: the same chain stack with the reference's own LotkaVolterraSolver
       (test/testSetup.py:101-141, scipy.solve_ivp; RK23 rtol 1e-2 coarse / DOP853 rtol 1e-5 fine,
       example_inference_lotkaVolterra_twoLevel.py:29-44),

one independent single-chain process per core (the reference is single-threaded and single-chain), its own
numpy RNG.  Writes profiles/r02_reference_python_timing.json, which bench.py copies into its JSON line as
`cpu_baseline.python_reference` -- clearly labelled as measured in the build container, not on the GPU box
(the reference cannot travel there).

    python oracle/time_reference.py [steps_per_chain]
"""
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np
from numpy.random import Generator, Philox

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_harness as rh          # noqa: E402
import make_golden as mg          # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "r02_reference_python_timing.json")


def _worker(args):
    kind, seed, nSteps, burn = args
    import numpy.random as npr
    from yagremcmc.postprocessing.autocorrelation import integrated_autocorrelation
    npr.seed(seed)
    p = mg.lv_problem()
    data = rh.Data(p['data'])
    noise = rh.CentredGaussianNoise(rh.IIDCovarianceMatrix(2, p['noiseVar']))
    prior = rh.Gaussian(rh.LotkaVolterraParameter.from_coefficient(p['priorMean']), rh.IIDCovarianceMatrix(2, p['priorVar']))
    if kind == 'rk4':
        solC = rh.RK4LotkaVolterraSolver(p['design'], dict(p['cfg'], rk4Steps=p['Nc']))
        solF = rh.RK4LotkaVolterraSolver(p['design'], dict(p['cfg'], rk4Steps=p['Nf']))
    else:
        from yagremcmc.test.testSetup import LotkaVolterraSolver
        solC = LotkaVolterraSolver(p['design'], dict(p['cfg'], solver='RK23', rtol=1e-2))
        solF = LotkaVolterraSolver(p['design'], dict(p['cfg'], solver='DOP853', rtol=1e-5))
    likC = rh.AdditiveGaussianNoiseLikelihood(data, rh.ForwardModel(solC), noise)
    likF = rh.AdditiveGaussianNoiseLikelihood(data, rh.ForwardModel(solF), noise)
    b = rh.MLDABuilder()
    b.bayesModel = rh.BayesianRegressionModelHierarchy(rh.Hierarchy([likC, likF]), rh.SharedComponent(prior, 2))
    b.baseProposalCovariance = rh.IIDCovarianceMatrix(2, 0.1)
    b.subChainLengths = [3]
    mcmc = rh.quiet(b.build_method)
    init = rh.LotkaVolterraParameter.from_coefficient(p['truth'] + 0.05 * Generator(Philox(7000 + seed)).standard_normal(2))
    t0 = time.perf_counter()
    rh.quiet(mcmc.run, nSteps, init, verbose=False)
    dt = time.perf_counter() - t0
    traj = np.array([np.asarray(s, dtype=np.float64).reshape(-1) for s in mcmc.chain.trajectory])
    iat = int(integrated_autocorrelation(traj[burn:], 'max'))
    return dt, (nSteps - burn) // max(iat, 1), iat, mcmc.diagnostics.global_acceptance_rate()


def measure(kind, nProc, nSteps, burn):
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(nProc) as pool:
        res = pool.map(_worker, [(kind, 300 + c, nSteps, burn) for c in range(nProc)])
    wall = time.perf_counter() - t0
    dts = np.array([r[0] for r in res])
    return dict(processes=nProc, steps_per_chain=nSteps, burn_in=burn,
                chain_steps_per_s=float(nProc * (nSteps - 1) / dts.max()),
                chain_steps_per_s_per_core=float(np.mean((nSteps - 1) / dts)),
                ess_per_s=float(sum(r[1] for r in res) / dts.max()),
                mean_iat_max=float(np.mean([r[2] for r in res])), acceptance=float(np.mean([r[3] for r in res])),
                wall_s=float(wall))


if __name__ == "__main__":
    nSteps = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    nProc = len(os.sched_getaffinity(0))
    import platform
    out = dict(
        what="UNMODIFIED rkutri/yagre-mcmc (pure Python) on the C5 problem, one single-chain process per core; "
             "measured in the BUILD CONTAINER (not on the GPU box: the reference cannot travel there)",
        host=dict(cores=nProc, machine=platform.processor() or platform.machine(), python=platform.python_version(),
                  numpy=np.__version__),
        definition="chain_steps_per_s = processes x transitions / slowest process; ess_per_s = sum_chains (N - burnIn) // "
                   "IAT_max / slowest process (example_inference_lotkaVolterra_twoLevel.py:117-118,132)",
        rk4_plugin=measure('rk4', nProc, nSteps, 100),
        as_shipped_solve_ivp=measure('ivp', nProc, nSteps, 100))
    json.dump(out, open(OUT, "w"), indent=1)
    print(json.dumps(out, indent=1))
