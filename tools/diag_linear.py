"""Dev tool: convergence of the C3 linear ensemble mean towards the closed-form posterior."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench_problems as bp
from yagre_mcmc_b200.ensemble import ChainEnsemble, LoweredProblem
nc = int(os.environ.get("NCH", 262144))
two = os.environ.get("LEVELS", "2") == "2"
meta, arrays = bp.linear_problem(two)
mean, cov = bp.linear_posterior('f')
sd = np.sqrt(np.diag(cov))
print("posterior mean", mean, "std", sd, "expected SE", sd / np.sqrt(nc))
ens = ChainEnsemble(LoweredProblem(meta, arrays), nc, seed=int(os.environ.get("SEED", 17)))
ens.set_state(np.tile(mean, (nc, 1)))
ens.run(int(os.environ.get("BURN", 60000)), samples=False)
zs = []
for k in range(int(os.environ.get("SNAPS", 12))):
    ens.run(int(os.environ.get("GAP", 20000)), samples=False)
    th = ens.state()["theta"].cpu().numpy()
    z = (th.mean(1) - mean) / (sd / np.sqrt(nc))
    zs.append(z)
    print(k, "z-score of mean", z, "var ratio", th.var(1) / np.diag(cov), flush=True)
zs = np.array(zs)
print("rms z", np.sqrt((zs ** 2).mean(0)), "mean z", zs.mean(0))
