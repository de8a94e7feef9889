import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np, torch
from golden_io import load
from yagre_mcmc_b200.ensemble import ChainEnsemble, LoweredProblem
meta,a = load("mlda3_gauss2d")
pb = LoweredProblem(meta,a)
nc, ns = a["u_f"].shape
ens = ChainEnsemble(pb, nc)
ens.set_state(a["theta0"])
torch.cuda.synchronize(); print("set_state ok")
z = np.ascontiguousarray(np.transpose(a["z"], (1, 2, 3, 0))); u_c = np.ascontiguousarray(np.transpose(a["u_c"], (1, 2, 0))); u_f = np.ascontiguousarray(np.transpose(a["u_f"], (1, 0)))
out = ens.run(ns, samples=True, accepted=True, logpost=True, inject=dict(z=z,u_c=u_c,u_f=u_f))
torch.cuda.synchronize(); print("inject run ok")
acc = out["accepted"].cpu().numpy().T
print("flips", int((acc != a["accepted"]).sum()))
ens2 = ChainEnsemble(pb, 1000, seed=3); ens2.set_state(np.tile([-8.,-7.],(1000,1)))
out = ens2.run(50, samples=True); torch.cuda.synchronize(); print("philox WS run ok", ens2.last_launch())
