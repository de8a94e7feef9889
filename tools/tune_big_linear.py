"""Dev tool: throughput of the DMMA linear-model kernel against the measured FP64 tensor / vector peaks."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench_problems as bp
from yagre_mcmc_b200.ensemble import ChainEnsemble, LoweredProblem, fp64_peak_tflops, fp64_tensor_peak_tflops

print("fp64 vector peak", fp64_peak_tflops(0, 30.0), "fp64 tensor (DMMA) peak", fp64_tensor_peak_tflops(0, 30.0), flush=True)
cases = [(64, 256, 1, False, 1), (32, 256, 1, False, 1), (16, 128, 1, False, 1), (64, 128, 1, True, 2), (32, 96, 2, True, 3)]
for nc in [int(x) for x in os.environ.get("NCH", "65536,16384").split(",")]:
    for d, dd, nd, two, J in cases:
        meta, arrays = bp.big_linear_problem(d, dd, nd, two_level=two, J=J)
        ens = ChainEnsemble(LoweredProblem(meta, arrays), nc, seed=1)
        mean, _ = bp.linear_gaussian_posterior(arrays, meta["levels"] - 1)
        ens.set_state(np.tile(mean, (nc, 1)))
        S = 50
        ens.run(S, samples=False)
        torch.cuda.synchronize()
        c0 = ens.counters()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            ens.run(S, samples=False)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        c1 = ens.counters()
        ev = (c1["coarse_evals"] - c0["coarse_evals"]) + (c1["fine_evals"] - c0["fine_evals"])
        flops = bp.big_linear_flops_per_eval(d, dd) * ev
        steps = nc * S * 3
        print(json.dumps(dict(chains=nc, d=d, data_dim=dd, n_data=nd, levels=2 if two else 1, J=J, ms=round(ms, 3),
                              steps_per_s=steps / ms * 1e3, tflops=flops / ms * 1e-9,
                              acc=(c1["accepted"] - c0["accepted"]) / steps, launch=ens.last_launch())), flush=True)
        ens.close()
