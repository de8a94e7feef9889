"""Dev tool: sweep launch geometry of the LV two-level kernel and report chain-steps/s
and the fraction of the measured DFMA peak."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench_problems as bp
from yagre_mcmc_b200.ensemble import ChainEnsemble, LoweredProblem, fp64_peak_tflops

n_chains = int(os.environ.get("NCH", 65536))
S = int(os.environ.get("S", 40))
two = os.environ.get("LEVELS", "2") == "2"
meta, arrays = bp.lv_problem(two_level=two, Nc=int(os.environ.get('NC', 64)), Nf=int(os.environ.get('NF', 512)), J=int(os.environ.get('JJ', 3)))
pb = LoweredProblem(meta, arrays)
peak = fp64_peak_tflops(0, 30.0)
print("fp64 DFMA peak TFLOP/s:", peak, flush=True)
th0 = bp.lv_initial_states(n_chains)
configs = [tuple(int(x) for x in c.split("x")) for c in os.environ.get("CFGS", "4x256x32,4x256x16,4x256x64,2x512x32,2x512x16,2x256x32,3x256x32,8x128x32,1x512x16,4x256x8,3x320x32").split(",")]
for cfg_ in configs:
    bps, thr, seg = cfg_[:3]
    grp = cfg_[3] if len(cfg_) > 3 else 0
    ens = ChainEnsemble(pb, n_chains, seed=1, blocks_per_sm=bps, threads_per_block=thr, rk4_segment=seg)
    ens.set_state(th0)
    ens.run(int(os.environ.get('BURN', 100)), samples=False)           # burn-in / warm-up
    torch.cuda.synchronize()
    c0 = ens.counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ens.run(S, samples=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    c1 = ens.counters()
    ce, fe = c1['coarse_evals'] - c0['coarse_evals'], c1['fine_evals'] - c0['fine_evals']
    nd = meta['n_data']
    flops = bp.lv_flops_per_eval(nd, meta['Nc']) * ce + bp.lv_flops_per_eval(nd, meta['Nf']) * fe if two else bp.lv_flops_per_eval(nd, meta['Nf']) * fe
    steps = n_chains * S * 3
    acc = (c1['accepted'] - c0['accepted']) / steps
    print(json.dumps(dict(bps=bps, thr=thr, seg=seg, grp=grp, frac=round(flops / ms * 1e-9 / peak, 4), launch=ens.last_launch(), ms=ms, steps_per_s=steps / ms * 1e3,
                          tflops=flops / ms * 1e-9, acc=acc,
                          fine_frac=fe / steps, coarse_per_step=ce / steps)), flush=True)
    ens.close()
