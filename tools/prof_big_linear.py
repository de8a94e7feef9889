"""Dev tool: one short run of the DMMA linear kernel (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench_problems as bp
from yagre_mcmc_b200.ensemble import ChainEnsemble, LoweredProblem
nc = 65536
# SHAPE = "d,data_dim,n_data,levels,J" (default: the 64 x 256 single-level case)
d, dd, nd, lv, J = (int(x) for x in os.environ.get("SHAPE", "64,256,1,1,1").split(","))
meta, arrays = bp.big_linear_problem(d, dd, nd, two_level=lv == 2, J=J)
ens = ChainEnsemble(LoweredProblem(meta, arrays), nc, seed=1, welford=os.environ.get("WELFORD", "1") == "1")
mean, _ = bp.linear_gaussian_posterior(arrays, meta["levels"] - 1)
ens.set_state(np.tile(mean, (nc, 1)))
for _ in range(3):
    ens.run(int(os.environ.get("STEPS", "10")), samples=False)
torch.cuda.synchronize()
print("ok", ens.counters())
