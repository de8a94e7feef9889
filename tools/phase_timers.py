"""Dev tool: per-CTA phase timers of the LV kernel (library built with -DYG_TIMERS)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from yagre_mcmc_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libyagre_b200_timers.so")
import bench_problems as bp
from yagre_mcmc_b200.ensemble import ChainEnsemble, LoweredProblem
meta, arrays = bp.lv_problem(True)
pb = LoweredProblem(meta, arrays)
n_chains, S = 65536, 20
for bps, thr, seg in [(1, 1024, 64), (1, 1024, 128), (4, 256, 64), (2, 512, 64)]:
    ens = ChainEnsemble(pb, n_chains, seed=1, blocks_per_sm=bps, threads_per_block=thr, rk4_segment=seg)
    ens.set_state(bp.lv_initial_states(n_chains))
    ens.run(100, samples=False)
    out = ens.run(S, samples=True)
    torch.cuda.synchronize()
    grid = ens.last_launch()['grid']
    t = out['samples'].view(torch.int64).flatten()[:8 * grid].cpu().numpy().reshape(grid, 8).astype(np.float64)
    owner, coarse, fine, total, nc_, nf_ = [t[:, i] for i in range(6)]
    # ideal FP64-pipe cycles per CTA if it had 1/bps of an SM: instr * 2 cycles / (4 SMSP) * bps share
    warp_instr_c = nc_ * 10 * 64 * 30 / 32.0
    warp_instr_f = nf_ * 10 * 512 * 30 / 32.0
    ideal_c = warp_instr_c * 2 / 4 * bps
    ideal_f = warp_instr_f * 2 / 4 * bps
    print(f"bps={bps} thr={thr} seg={seg}: per CTA mean cycles total={total.mean():.3e} (min {total.min():.3e} max {total.max():.3e}) "
          f"owner={owner.mean()/total.mean():.3f} coarse={coarse.mean()/total.mean():.3f} fine={fine.mean()/total.mean():.3f} | "
          f"coarse eff={ideal_c.mean()/coarse.mean():.3f} fine eff={ideal_f.mean()/fine.mean():.3f} "
          f"overall eff={(ideal_c+ideal_f).mean()/total.max():.3f}")
    ens.close()
