"""Dev tool: short runs of the one-chain-per-thread kernels (C3 linear two level, C2 adaptive Gaussian) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench_problems as bp
from yagre_mcmc_b200.ensemble import ChainEnsemble, LoweredProblem
meta, arrays = bp.linear_problem(True)
ens = ChainEnsemble(LoweredProblem(meta, arrays), 16384, seed=5)
ens.set_state(np.zeros((16384, 2)))
for _ in range(3):
    ens.run(2000, samples=False)
meta, arrays = bp.gauss2d_problem()
e2 = ChainEnsemble(LoweredProblem(meta, arrays), 4096, seed=5, adaptive=dict(idle=100, collection=200, eps=1e-4))
e2.set_state(np.tile([-8.0, -7.0], (4096, 1)))
for _ in range(3):
    e2.run(2000, samples=False)
torch.cuda.synchronize()
print("ok")
