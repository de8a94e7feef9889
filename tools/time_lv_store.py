"""Dev tool: cost of the sample write-back of lv_mh_kernel (C5, 65,536 chains)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench_problems as bp
from yagre_mcmc_b200.ensemble import ChainEnsemble, LoweredProblem
meta, arrays = bp.lv_problem(True)
nc = 65536
ens = ChainEnsemble(LoweredProblem(meta, arrays), nc, seed=1)
ens.set_state(bp.lv_initial_states(nc))
ens.run(100, samples=False)
buf = torch.empty((1000, 2, nc), dtype=torch.float64, device="cuda")
acc = None
def t(label, **kw):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ens.run(1000, **kw); e1.record(); torch.cuda.synchronize()
    print(f"{label:40s} {e0.elapsed_time(e1) / 1000:.5f} ms per transition", flush=True)
for rep in range(2):
    t("no outputs", samples=False)
    t("samples (preallocated)", samples_out=buf)
    t("samples thin=10", samples=True, thin=10)
    t("samples + accepted + logpost", samples=True, accepted=True, logpost=True)
    t("accepted only", samples=False, accepted=True)
print(ens.last_launch())
