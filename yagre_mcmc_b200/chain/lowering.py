"""Lowering: the object graph a ChainBuilder holds -> the plain arrays of yg_problem.

This is the host half of the drop-in boundary: everything the reference evaluates lazily per
step through Python objects (chain/builder.py:72-83, chain/method/mrw.py:75-86,
chain/method/mlda.py:315-343) is resolved ONCE here."""
import numpy as np

from ..ensemble import LoweredProblem
from .target import UnnormalisedPosterior


def _lower_density(density):
    """-> (model name, level dict) for one target density."""
    if isinstance(density, UnnormalisedPosterior):
        return density.device_level()
    return 'gauss', density.device_target()          # raises NotImplementedError when not Gaussian


def lower_problem(targets, proposalCov, subChainLength=None, equality='exact', proposal='mrw', pcnStep=None,
                  pcnMean=None):
    """targets: [target] (MRW), [surrogate, target] (two-level delayed acceptance) or
    [base surrogate, finest surrogate, target] (MLDA with two surrogates as the reference runs it,
    mlda.py:12-43,60-71,112-117: sub-chain on the base surrogate, screen with the finest one)."""
    if len(targets) not in (1, 2, 3):
        raise NotImplementedError(
            f"{len(targets)} levels requested: the device implements MRW, two-level delayed acceptance and MLDA "
            "with two surrogates; the reference itself crashes for 3 and 5 surrogates "
            "(AttributeError in SurrogateTransition, mlda.py:23-33)")
    lowered = [_lower_density(t) for t in targets]
    models = {m for m, _ in lowered}
    if len(models) != 1:
        raise NotImplementedError(f"levels of different model kinds {sorted(models)} cannot share one kernel")
    model = models.pop()
    arrays = {'prop_L': np.asarray(proposalCov.chol_factor(), dtype=np.float64)}
    dim = arrays['prop_L'].shape[0]
    for l, (_, lvl) in enumerate(lowered):
        for k, v in lvl.items():
            arrays[f"L{l}_{k}"] = np.asarray(v, dtype=np.float64)
        key = f"L{l}_g_mean" if model == 'gauss' else f"L{l}_prior_mean"
        if arrays[key].size != dim:
            raise ValueError(f"level {l}: parameter dimension {arrays[key].size} does not match the "
                             f"proposal covariance ({dim})")
    meta = dict(model=model, dim=dim, levels=len(targets), J=int(subChainLength or 1), eq=equality)
    if proposal == 'pcn':
        if len(targets) != 1:
            raise NotImplementedError("pCN is a single-level method (chain/method/pcn.py)")
        meta.update(proposal='pcn', pcn_step=float(pcnStep))
        arrays['pcn_mean'] = np.zeros(dim) if pcnMean is None else np.asarray(pcnMean, dtype=np.float64)
    return LoweredProblem(meta, arrays)
