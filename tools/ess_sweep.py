"""Dev tool: ESS/s of C5 over the sub-chain length J and the scale of the pooled proposal covariance (16,384 chains,
2,000 transitions after a 400-transition burn-in that is charged to the time; IAT on the device).  The example's setting
is J = 3 with a fixed 0.1 I proposal."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench_problems as bp
from yagre_mcmc_b200.ensemble import ChainEnsemble, LoweredProblem, iat_ess
from yagre_mcmc_b200.parallel import pooled_diagnostics, proposal_factor_from_covariance

nc, ns, burn = 16384, 2000, 400
th0 = bp.lv_initial_states(nc)
meta, arrays = bp.lv_problem(True)
ens = ChainEnsemble(LoweredProblem(meta, arrays), nc, seed=3)
ens.set_state(th0)
ens.run(600, samples=False)
cov = pooled_diagnostics(ens)["covariance"]
start = ens.state()["theta"].t().contiguous().cpu().numpy()
ens.close()
print("pooled posterior covariance", cov.tolist(), flush=True)
rows = []
for J in (1, 2, 3, 4, 6, 8, 12):
    for scale in (None, 0.36, 0.72, 1.44, 2.88, 5.76):
        m, a = bp.lv_problem(True, J=J)
        if scale is not None:
            a = dict(a); a["prop_L"] = proposal_factor_from_covariance(cov, scale, 1e-10)
        e = ChainEnsemble(LoweredProblem(m, a), nc, seed=4)
        e.set_state(start)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        smp = e.run(burn + ns, samples=True)["samples"]
        e1.record(); torch.cuda.synchronize()
        iat, ess = iat_ess(smp[burn:], "max")
        c = e.counters()
        rows.append(dict(J=J, scale="0.1 I (example)" if scale is None else scale, ess_per_s=float(ess.sum().item()) / (e0.elapsed_time(e1) * 1e-3),
                         iat=float(iat.double().mean().item()), accept=c["accepted"] / c["transitions"],
                         fine_evals_per_step=c["fine_evals"] / c["transitions"], steps_per_s=nc * (burn + ns) / (e0.elapsed_time(e1) * 1e-3)))
        print(json.dumps(rows[-1]), flush=True)
        del smp
        e.close()
best = max(rows, key=lambda r: r["ess_per_s"])
ref = [r for r in rows if r["J"] == 3 and r["scale"] == "0.1 I (example)"][0]
print("example:", json.dumps(ref)); print("best:", json.dumps(best), "ratio", best["ess_per_s"] / ref["ess_per_s"])
