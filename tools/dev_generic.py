"""Dev tool: the one-chain-per-thread kernels on small ensembles (C2, C3): plain Philox mode (warp-specialised
variant for <= 32,768 chains) against record mode (unspecialised kernel) bit for bit, and chain-steps/s.
LIB=dev uses tools/_build/libyagre_b200_dev.so (make -C yagre_mcmc_b200/csrc dev)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from yagre_mcmc_b200 import _lib
if os.environ.get("LIB") == "dev":
    _lib.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libyagre_b200_dev.so")
import bench_problems as bp
from yagre_mcmc_b200.ensemble import ChainEnsemble, LoweredProblem

AM = dict(idle=100, collection=200, eps=1e-4)
cases = [("C3 linear two level", bp.linear_problem(True), lambda n: np.zeros((n, 2)), None, 16384, 5000),
         ("C3 linear single level", bp.linear_problem(False), lambda n: np.zeros((n, 2)), None, 16384, 5000),
         ("C2 gauss2d adaptive", bp.gauss2d_problem(), lambda n: np.tile([-8.0, -7.0], (n, 1)), AM, 4096, 20000),
         ("C2 gauss2d", bp.gauss2d_problem(), lambda n: np.tile([-8.0, -7.0], (n, 1)), None, 4096, 20000),
         ("gauss2d 1 chain", bp.gauss2d_problem(), lambda n: np.tile([-8.0, -7.0], (n, 1)), None, 1, 20000),
         ("gauss2d 65536 chains (plain kernel)", bp.gauss2d_problem(), lambda n: np.tile([-8.0, -7.0], (n, 1)), None, 65536, 5000)]
for name, (meta, arrays), init, am, nc, S in cases:
    pb = LoweredProblem(meta, arrays)
    # bitwise: plain mode vs record mode, same seed (333 chains: a ragged last warp)
    n_small = 1 if nc == 1 else 333
    outs = []
    for rec in (False, True):
        e = ChainEnsemble(pb, n_small, seed=77, adaptive=am)
        e.set_state(init(n_small))
        o = e.run(400, samples=True, accepted=True, record=rec)
        outs.append((o["samples"].clone(), o["accepted"].clone(), e.counters(), e.state()))
        e.close()
    same = torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    same = same and outs[0][2] == outs[1][2] and all(torch.equal(outs[0][3][k], outs[1][3][k]) for k in outs[0][3] if torch.is_tensor(outs[0][3][k]))
    ens = ChainEnsemble(pb, nc, seed=5, adaptive=am)
    ens.set_state(init(nc))
    ens.run(S, samples=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ens.run(S, samples=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{name:38s} bitwise plain==record: {same}   {nc * S * 3 / ms * 1e3:.4e} chain-steps/s   launch {ens.last_launch()}", flush=True)
    ens.close()
