"""Dev tool: per-warp phase timers of the LV kernel (library built with `make -C yagre_mcmc_b200/csrc timers`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from yagre_mcmc_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libyagre_b200_timers.so")
import bench_problems as bp
from yagre_mcmc_b200.ensemble import ChainEnsemble, LoweredProblem
two = os.environ.get("LEVELS", "2") == "2"
Nc, Nf = int(os.environ.get("NC", 64)), int(os.environ.get("NF", 512))
meta, arrays = bp.lv_problem(two, Nc=Nc, Nf=Nf)
pb = LoweredProblem(meta, arrays)
n_chains, S = 65536, 20
names = ["owners", "wait1", "noise", "eval", "wait2", "commit"]
for bps, thr, seg in [(1, int(os.environ.get("THR", 768)), 128)]:
    ens = ChainEnsemble(pb, n_chains, seed=1, blocks_per_sm=bps, threads_per_block=thr, rk4_segment=seg)
    ens.set_state(bp.lv_initial_states(n_chains))
    ens.run(100, samples=False)
    if os.environ.get("SKIP"):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ens.run(21, samples=False); e1.record(); torch.cuda.synchronize()
        t_full = e0.elapsed_time(e1)
        e0.record(); ens.run(21, thin=7, samples=False); e1.record(); torch.cuda.synchronize()
        t_skip = e0.elapsed_time(e1)
        e0.record(); ens.run(21, thin=7, samples=False); e1.record(); torch.cuda.synchronize()
        t_skip2 = e0.elapsed_time(e1)
        print(f"21 transitions: full {t_full:.3f} ms, without integration {t_skip:.3f} / {t_skip2:.3f} ms -> {t_skip2 / t_full:.4f} of the launch")
    out = ens.run(S, samples=True)
    torch.cuda.synchronize()
    grid = ens.last_launch()["grid"]
    nw = thr // 32
    t = out["samples"].view(torch.int64).flatten()[:10 * 32 * grid].cpu().numpy().reshape(grid, 32, 10)[:, :nw].astype(np.float64)   # the kernel reserves 32 warp slots per CTA
    total = t[:, :, 8]
    print(f"bps={bps} thr={thr} seg={seg}: CTA total cycles mean {total.mean():.4e} min {total.min():.4e} max {total.max():.4e}")
    for k, nm in enumerate(names):
        f = t[:, :, k] / total
        print(f"  {nm:7s} share of CTA time: mean over warps {f.mean():.4f}  first half {f[:, :nw // 2].mean():.4f}  second half {f[:, nw // 2:].mean():.4f}  min {f.min():.4f} max {f.max():.4f}")
    items_c, items_f = t[:, 0, 6], t[:, 0, 7]
    if not two:
        items_c, items_f = np.zeros_like(items_c), items_c
    ideal = (items_c * Nc + items_f * Nf) * 23 / 32.0 * 2 / 4 * bps     # 23 issue slots per RK4 step
    print(f"  ideal FP64-pipe cycles / CTA total: mean {np.mean(ideal / total[:, 0]):.4f}; vs slowest CTA {ideal.mean() / total.max():.4f}")
    ens.close()
